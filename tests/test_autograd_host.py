"""Autograd through diag (PyTorch engine) against central finite differences of the NumPy-engine
path -- the way the reference's tests/test_grad.py::function_grad_test checks it.  CPU: kernels
emulated; the same test body runs on the GPU in tests/test_gpu_autograd.py."""
import numpy as np
import pytest
import torch

import sqcircuit_b200 as sq
from emul_kernels import EmulKernels


def build(kind, kernels, torch_engine, delta=None):
    sq.set_engine('PyTorch' if torch_engine else 'NumPy')
    rg = torch_engine
    scale = {} if delta is None else delta
    if kind == 'transmon':
        C = sq.Capacitor(7.746 * scale.get('C', 1), 'fF', requires_grad=rg)
        J = sq.Junction(5 * scale.get('J', 1), 'GHz', requires_grad=rg)
        cr = sq.Circuit({(0, 1): [C, J]}, kernels=kernels)
        cr.set_trunc_nums([40])
        els = {'C': C, 'J': J}
    elif kind == 'fluxonium':
        lp = sq.Loop(0.3 * scale.get('loop', 1), requires_grad=rg)
        C = sq.Capacitor(3.6 * scale.get('C', 1), 'GHz', requires_grad=rg)
        L = sq.Inductor(0.46 * scale.get('L', 1), 'GHz', loops=[lp], requires_grad=rg)
        J = sq.Junction(10.2 * scale.get('J', 1), 'GHz', loops=[lp], requires_grad=rg)
        cr = sq.Circuit({(0, 1): [C, L, J]}, flux_dist='junctions', kernels=kernels)
        cr.set_trunc_nums([60])
        els = {'C': C, 'L': L, 'J': J, 'loop': lp}
    elif kind == 'zeropi':
        lp = sq.Loop(0.31 * scale.get('loop', 1), requires_grad=rg)
        C = sq.Capacitor(0.15 * scale.get('C', 1), 'GHz', requires_grad=rg)
        CJ = sq.Capacitor(10 * scale.get('CJ', 1), 'GHz', requires_grad=rg)
        J = sq.Junction(5 * scale.get('J', 1), 'GHz', loops=[lp], requires_grad=rg)
        L = sq.Inductor(0.13 * scale.get('L', 1), 'GHz', loops=[lp], requires_grad=rg)
        cr = sq.Circuit({(0, 1): [CJ, J], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L], (2, 3): [CJ, J]},
                        kernels=kernels)
        # the PyTorch engine keeps the uncoupled harmonic mode (truncated to one level here)
        cr.set_trunc_nums([44, 1, 9] if torch_engine else [44, 9])
        els = {'C': C, 'CJ': CJ, 'J': J, 'L': L, 'loop': lp}
    else:
        raise ValueError(kind)
    return cr, els


def grad_check(kind, kernels, n_eig=6, which=(1, 0), rtol=2e-5):
    cr, els = build(kind, kernels, True)
    ef, ev = cr.diag(n_eig)
    f = ef[which[0]] - ef[which[1]]
    f.backward()
    grads = {k: float(e.grad) for k, e in els.items()}
    base = float(f.detach())
    sq.set_engine('NumPy')
    for name, g in grads.items():
        d = 1e-5
        vals = []
        for sgn in (+1, -1):
            crn, eln = build(kind, kernels, False, {name: 1 + sgn * d})
            e, _ = crn.diag(n_eig)
            vals.append(e[which[0]] - e[which[1]])
        crn, eln = build(kind, kernels, False)
        x0 = eln[name].internal_value if name != 'loop' else eln[name].internal_value
        if name in ('C', 'L', 'CJ'):
            # GHz-valued elements: internal value ~ 1/energy -> d(internal)/d(scale) = -x0
            dx = -2 * d * x0
            if kind == 'transmon' and name == 'C':
                dx = 2 * d * x0          # given in fF: internal value scales with the input
        else:
            dx = 2 * d * x0
        fd = (vals[0] - vals[1]) / dx
        assert abs(g - fd) <= rtol * abs(fd) + 1e-12 * abs(base / x0), (name, g, fd)
    sq.set_engine('NumPy')


def test_autograd_transmon_cpu():
    grad_check('transmon', EmulKernels())


def test_autograd_fluxonium_cpu():
    grad_check('fluxonium', EmulKernels())


def test_autograd_zeropi_three_modes_cpu():
    # Hellmann-Feynman with fixed Fock operators is exact only for a converged harmonic truncation:
    # same 2e-2 tolerance as the reference's tests/test_grad.py
    grad_check('zeropi', EmulKernels(), n_eig=5, rtol=2e-2)


def eigvec_grad_check(kernels):
    """tests/test_papers.py::test_paper_2_a: g_10 = |<1|Q|0>| / C, gradient w.r.t. C."""
    def g10(cr, C, torch_engine):
        ef, ev = cr.diag(20)
        Q = cr.charge_op(1, 'original')
        if torch_engine:
            return 1 / C.get_value() * torch.abs(ev[1].conj() @ Q @ ev[0])
        v = [np.asarray(x)[:, 0] for x in ev]
        return 1 / C.get_value() * np.abs(v[1].conj() @ Q @ v[0])
    sq.set_engine('PyTorch')
    C = sq.Capacitor(12, 'fF', requires_grad=True)
    cr = sq.Circuit({(0, 1): [C, sq.Junction(10, 'GHz')]}, kernels=kernels)
    cr.set_trunc_nums([40])
    val = g10(cr, C, True)
    val.backward()
    grad = float(C.grad)
    sq.set_engine('NumPy')
    vals = []
    d = 1e-6
    for sgn in (+1, -1):
        Cn = sq.Capacitor(12 * (1 + sgn * d), 'fF')
        crn = sq.Circuit({(0, 1): [Cn, sq.Junction(10, 'GHz')]}, kernels=kernels)
        crn.set_trunc_nums([40])
        vals.append(g10(crn, Cn, False))
    fd = (vals[0] - vals[1]) / (2 * d * 12e-15)
    assert abs(grad - fd) <= 2e-2 * abs(fd), (grad, fd)
    assert np.isclose(float(val.detach()), 0.5 * (vals[0] + vals[1]), rtol=1e-6)


def test_autograd_eigenvector_path_cpu():
    eigvec_grad_check(EmulKernels())
