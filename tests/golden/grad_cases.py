"""Circuits of the gradient fixtures (tests/golden/gradients.npz).  The same constructor calls build them from
the reference package (make_golden.py, in the container that has /root/reference) and from the product package
(the tests)."""
import numpy as np


def grad_circuits(sqm, kind, kernels=None, rg=True):
    """Circuits of the gradient fixtures, built with module ``sqm`` (the reference here, the product package in
    the tests -- same constructor calls).  Returns (circuit, {name: element}, trunc, n_eig).
    rg: elements require gradients (PyTorch engine); False builds the same circuit for the NumPy engine."""
    kw = {} if kernels is None else {'kernels': kernels}
    if kind == 'fluxonium':
        lp = sqm.Loop(0.3, requires_grad=rg)
        C = sqm.Capacitor(3.6, 'GHz', Q=1e6, requires_grad=rg)
        L = sqm.Inductor(0.46, 'GHz', Q=500e6, loops=[lp], requires_grad=rg)
        J = sqm.Junction(10.2, 'GHz', A=1e-7, x=3e-06, loops=[lp], requires_grad=rg)
        cr = sqm.Circuit({(0, 1): [C, L, J]}, flux_dist='junctions', **kw)
        return cr, {'C': C, 'L': L, 'J': J, 'loop': lp}, [60], 8
    if kind == 'flux_transmon':
        lp = sqm.Loop(0.2, requires_grad=rg)
        C = sqm.Capacitor(7.746, 'fF', Q=1e6, requires_grad=rg)
        J1 = sqm.Junction(5, 'GHz', loops=[lp], requires_grad=rg)
        J2 = sqm.Junction(5.5, 'GHz', loops=[lp], requires_grad=rg)
        cr = sqm.Circuit({(0, 1): [C, J1, J2]}, **kw)
        return cr, {'C': C, 'J1': J1, 'J2': J2, 'loop': lp}, [20], 6
    if kind == 'zeropi':
        lp = sqm.Loop(0.31, requires_grad=rg)
        C = sqm.Capacitor(0.15, 'GHz', requires_grad=rg)
        CJ = sqm.Capacitor(10, 'GHz', requires_grad=rg)
        J = sqm.Junction(5, 'GHz', loops=[lp], requires_grad=rg)
        L = sqm.Inductor(0.13, 'GHz', loops=[lp], requires_grad=rg)
        cr = sqm.Circuit({(0, 1): [CJ, J], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, J]}, **kw)
        return cr, {'C': C, 'CJ': CJ, 'J': J, 'L': L, 'loop': lp}, [30, 1, 9], 6
    if kind.startswith('discovery'):
        # BASELINE configs[4]: 4-mode circuit (two harmonic, two charge modes), candidate number in the name
        idx = int(kind[len('discovery'):] or 0)
        v = discovery_values(idx)
        lp = sqm.Loop(0.37, requires_grad=rg)
        names = ['C01', 'J01', 'C12', 'L12', 'C20', 'J20', 'C23', 'C30', 'J30', 'C34', 'C40', 'L40']
        mk = {'C': sqm.Capacitor, 'J': sqm.Junction, 'L': sqm.Inductor}
        els = {}
        for nm, val in zip(names, v):
            extra = {'loops': [lp]} if nm in ('J01', 'L12', 'J20') else {}
            els[nm] = mk[nm[0]](float(val), 'GHz', requires_grad=rg, **extra)
        cr = sqm.Circuit({(0, 1): [els['C01'], els['J01']], (1, 2): [els['C12'], els['L12']],
                          (2, 0): [els['C20'], els['J20']], (2, 3): [els['C23']],
                          (3, 0): [els['C30'], els['J30']], (3, 4): [els['C34']],
                          (4, 0): [els['C40'], els['L40']]}, **kw)
        els['loop'] = lp
        return cr, els, [5, 5, 3, 3], 6
    raise ValueError(kind)


def discovery_values(idx):
    """Element values (GHz) of candidate ``idx`` of the 4-mode discovery circuit: a fixed base design times a
    deterministic +-20 % spread."""
    base = np.array([3.0, 8.0, 5.0, 0.6, 4.0, 6.0, 20.0, 1.2, 12.0, 25.0, 0.8, 1.5])
    rng = np.random.default_rng(1000 + idx)
    return base * (1 + 0.2 * rng.uniform(-1, 1, len(base))) if idx else base
