"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference)
in this container.  QuTiP/matplotlib/IPython are replaced by the stand-ins in
``ref_shims`` (see its README).  Run:  python tests/golden/make_golden.py

Outputs (committed): tests/golden/*.npz.  /root/reference does not exist on the
GPU box, so tests only ever read the .npz files.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import scipy.linalg  # noqa: E402
import SQcircuit as sq  # noqa: E402

DEC_ALL = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']


def exact_eigs(cr, k):
    """Dense LAPACK eigenpairs of the reference's own Hamiltonian matrix."""
    H = cr.hamiltonian().full()
    w, v = scipy.linalg.eigh(H, subset_by_index=[0, k - 1])
    return w, v


def transform_record(cr):
    return dict(S=cr.S, R=cr.R, wTrans=np.asarray(cr.wTrans, float), cInvTrans=cr.cInvTrans,
                lTrans=cr.lTrans, omega=cr.omega, B=np.asarray(cr.B, float),
                C=np.asarray(cr.C, float), L=np.asarray(cr.L, float))


def sweep_record(cr, set_param, params, k, decs, states=(1, 0), mel=None, store_vecs=()):
    """Per sweep point: reference diag() efreqs (ARPACK), exact efreqs of the reference H
    (LAPACK), decay rates and matrix elements from the reference with ITS eigenvectors."""
    ef_ref, ef_exact, dec, mels, vecs = [], [], {d: [] for d in decs}, [], {}
    for i, p in enumerate(params):
        set_param(p)
        e, _ = cr.diag(k)
        ef_ref.append(np.array(e))
        w, v = exact_eigs(cr, k)
        ef_exact.append(w / (2 * np.pi * 1e9))
        for d in decs:
            dec[d].append(float(np.real(cr.dec_rate(d, states))))
        if mel is not None:
            mels.append([complex(cr.matrix_elements(ct, nodes, st)) for ct, nodes, st in mel])
        if i in store_vecs:
            vecs['evecs_%d' % i] = v
    out = dict(params=np.array(params), efreqs_ref=np.array(ef_ref), efreqs_exact=np.array(ef_exact))
    for d in decs:
        out['dec_' + d] = np.array(dec[d])
    if mel is not None:
        out['mel_abs'] = np.abs(np.array(mels))
    out.update(vecs)
    return out


def zero_pi(trunc):
    loop1 = sq.Loop()
    C = sq.Capacitor(0.15, "GHz")
    CJ = sq.Capacitor(10, "GHz")
    JJ = sq.Junction(5, "GHz", loops=[loop1])
    L = sq.Inductor(0.13, "GHz", loops=[loop1])
    cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L],
                     (2, 3): [CJ, JJ]})
    cr.set_trunc_nums(trunc)
    return cr, loop1


def fluxonium(trunc):
    loop1 = sq.Loop()
    cap = sq.Capacitor(3.6, 'GHz', Q=1e6)
    ind = sq.Inductor(0.46, 'GHz', Q=500e6, loops=[loop1])
    junc = sq.Junction(10.2, 'GHz', cap=cap, A=1e-7, x=3e-06, loops=[loop1])
    cr = sq.Circuit({(0, 1): [ind, junc]}, flux_dist='all')
    cr.set_trunc_nums(trunc)
    return cr, loop1


def transmon(trunc):
    cr = sq.Circuit({(0, 1): [sq.Capacitor(20, 'GHz'), sq.Junction(15.0, 'GHz', A=1e-7)]})
    cr.set_trunc_nums(trunc)
    return cr


def transmon_single(ncut):
    """BASELINE configs[0]: JJ + capacitor, charge basis ncut=30."""
    cr = sq.Circuit({(0, 1): [sq.Capacitor(7.746, 'fF', Q=1e6), sq.Junction(5, 'GHz')]})
    cr.set_trunc_nums([ncut])
    return cr


def coupled_fluxonium_transmon(trunc):
    loop1 = sq.Loop()
    C = sq.Capacitor(1, 'GHz')
    L = sq.Inductor(1, 'GHz', loops=[loop1])
    JJ = sq.Junction(1, 'GHz', loops=[loop1])
    JJ2 = sq.Junction(1, 'GHz')
    C_c = sq.Capacitor(1, 'GHz')
    cr = sq.Circuit({(0, 1): [L, JJ, C], (0, 2): [JJ2, C], (1, 2): [C_c]})
    cr.set_trunc_nums(trunc)
    return cr, loop1


def trt(trunc):
    """BASELINE configs[3]: transmon - resonator - transmon (3 modes)."""
    Cq1, Cq2 = sq.Capacitor(0.30, 'GHz'), sq.Capacitor(0.27, 'GHz')
    J1, J2 = sq.Junction(14.0, 'GHz'), sq.Junction(17.0, 'GHz')
    Cr, Lr = sq.Capacitor(0.5, 'GHz'), sq.Inductor(8.0, 'GHz')
    Cc1, Cc2 = sq.Capacitor(12.0, 'GHz'), sq.Capacitor(15.0, 'GHz')
    # resonator on node 1: scipy>=1.17 sqrtm returns inf for diag(0, x, 0) (reference
    # circuit.py:913), so the singular susceptance matrix must not start with a zero.
    cr = sq.Circuit({(0, 1): [Cr, Lr], (0, 2): [Cq1, J1], (0, 3): [Cq2, J2], (1, 2): [Cc1],
                     (1, 3): [Cc2]})
    cr.set_trunc_nums(trunc)
    return cr


def inductively_shunted(trunc):
    loop1 = sq.Loop()
    C_r = sq.Capacitor(20.3, "fF")
    C_q = sq.Capacitor(5.3, "fF")
    L_r = sq.Inductor(15.6, "nH")
    L_q = sq.Inductor(386, "nH", loops=[loop1])
    L_s = sq.Inductor(4.5, "nH", loops=[loop1])
    JJ = sq.Junction(6.2, "GHz", loops=[loop1])
    cr = sq.Circuit({(0, 1): [C_r], (1, 2): [L_r], (0, 2): [L_s], (2, 3): [L_q], (0, 3): [JJ, C_q]})
    cr.set_trunc_nums(trunc)
    return cr, loop1


def main(only=None):
    sq.set_engine('NumPy')
    if only:
        return PARTS[only]()
    for f in PARTS.values():
        f()
    print('done')


def part_zeropi():
    # --- zero-pi (reference test truncation) : flux sweep, all loss channels -----------
    cr, lp = zero_pi([25, 13])
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lp.set_flux, np.linspace(0, 0.5, 9), 10, DEC_ALL,
                            mel=[('capacitive', (0, 1), (0, 1)), ('inductive', (0, 2), (0, 1)),
                                 ('capacitive', (2, 3), (1, 2))], store_vecs=(0, 3, 8)))
    lp.set_flux(0.31)
    rec['H_at_0p31'] = cr.hamiltonian().full()
    np.savez_compressed(os.path.join(HERE, 'zeropi_25_13.npz'), **rec)



def part_fluxonium():
    # --- fluxonium trunc=100 : BASELINE configs[1] -----------------------------------
    cr, lp = fluxonium([100])
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lp.set_flux, np.linspace(0, 1.0, 17), 10,
                            ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux'],
                            mel=[('capacitive', (0, 1), (0, 1)), ('inductive', (0, 1), (0, 1))],
                            store_vecs=(0, 4, 8)))
    lp.set_flux(0.31)
    rec['H_at_0p31'] = cr.hamiltonian().full()
    np.savez_compressed(os.path.join(HERE, 'fluxonium_100.npz'), **rec)



def part_transmon():
    # --- transmon (reference test) : charge sweep ------------------------------------
    cr = transmon([100])
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lambda ng: cr.set_charge_offset(1, float(ng)), np.linspace(0, 0.5, 5), 10,
                            ['capacitive', 'cc', 'charge'],
                            mel=[('capacitive', (0, 1), (0, 1))], store_vecs=(1,)))
    np.savez_compressed(os.path.join(HERE, 'transmon_100.npz'), **rec)

    # --- BASELINE configs[0]: single transmon ncut=30 ----------------------------------
    cr = transmon_single(30)
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lambda ng: cr.set_charge_offset(1, float(ng)), [0.0, 0.2], 10,
                            ['capacitive', 'cc', 'charge'], store_vecs=(0, 1)))
    rec['H_at_ng0p2'] = cr.hamiltonian().full()
    np.savez_compressed(os.path.join(HERE, 'transmon_ncut30.npz'), **rec)



def part_coupled():
    # --- coupled fluxonium-transmon (harmonic x charge, 2 modes) -----------------------
    cr, lp = coupled_fluxonium_transmon([30, 10])
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lp.set_flux, [0.0, 0.2, 0.5], 10, DEC_ALL, store_vecs=(1,)))
    np.savez_compressed(os.path.join(HERE, 'fluxonium_transmon_30_10.npz'), **rec)

    # --- inductively shunted (3 harmonic modes, one truncated to 1) ---------------------
    cr, lp = inductively_shunted([1, 9, 23])
    rec = transform_record(cr)
    rec.update(sweep_record(cr, lp.set_flux, [0.0, 0.25, 0.5], 6, ['capacitive', 'inductive', 'cc', 'flux']))
    np.savez_compressed(os.path.join(HERE, 'ind_shunted_1_9_23.npz'), **rec)



def part_trt():
    # --- transmon-resonator-transmon (BASELINE configs[3], small truncation) -----------
    cr = trt([6, 6, 6])
    rec = transform_record(cr)
    ngs = [(0.0, 0.0), (0.13, 0.41), (0.5, 0.25)]
    charge_modes = [i + 1 for i in range(cr.n) if cr._is_charge_mode(i)]

    def set_ng(p):
        for mode, v in zip(charge_modes, p):
            cr.set_charge_offset(mode, float(v))
    rec.update(sweep_record(cr, set_ng, ngs, 10, ['capacitive', 'cc', 'charge'], store_vecs=(1,)))
    rec['charge_modes'] = np.array(charge_modes)
    np.savez_compressed(os.path.join(HERE, 'trt_6_6_6.npz'), **rec)



# (Fock index, charge number) of the golden Hamiltonian columns at the bench truncation
H_COLS_100_51 = [(0, 0), (3, 1), (3, -1), (50, 25), (50, -25), (99, 50), (99, -50), (17, 7), (17, -7), (64, 0)]


def part_zeropi_big():
    # --- zero-pi at the bench truncation (N = 10100): a few exact eigenvalue sets --------
    cr, lp = zero_pi([100, 51])
    import scipy.sparse.linalg as spla
    fl = [0.0, 0.37, 0.5]
    ws = []
    for f in fl:
        lp.set_flux(f)
        H = cr.hamiltonian().data_as('csr_matrix')
        w = spla.eigsh(H, k=10, which='SA', tol=0, ncv=60)[0]
        ws.append(np.sort(w) / (2 * np.pi * 1e9))
    # columns of the reference Hamiltonian at flux 0.31 (Fock i, charge n pairs +-n), dense per column
    lp.set_flux(0.31)
    H = cr.hamiltonian().data_as('csr_matrix').tocsc()
    inner, nc = 101, 50
    cols = [i * inner + nc + n for i, n in H_COLS_100_51]
    Hc = np.stack([np.asarray(H[:, c].todense()).ravel() for c in cols], axis=1)
    np.savez_compressed(os.path.join(HERE, 'zeropi_100_51.npz'), params=np.array(fl),
                        efreqs_exact=np.array(ws), H_cols=np.array(cols), H_cols_at_0p31=Hc)



# ---- gradients from the reference's PyTorch engine (torch_extensions.py) -------------------------------------
from grad_cases import grad_circuits, discovery_values  # noqa: E402,F401  (shared with the tests)


def part_grad():
    """Values and gradients from the reference's PyTorch engine: transition frequency, every decoherence
    channel the circuit has (T1 channels through the eigenvector backward, dephasing channels through
    DecRateCharge / DecRateCC / DecRateFlux), all eigenfrequencies of the 4-mode circuits."""
    import torch
    from SQcircuit.settings import set_optim_mode
    set_optim_mode(True)
    rec = {}
    cases = {'fluxonium': ['capacitive', 'inductive', 'quasiparticle', 'cc', 'flux'],
             'flux_transmon': ['capacitive', 'charge', 'cc', 'flux'],
             'zeropi': ['inductive', 'charge', 'cc', 'flux']}   # (capacitive: the reference's sp_add fails on sparse tensors with this torch for n > 1)
    for kind, decs in cases.items():
        cr, els, trunc, k = grad_circuits(sq, kind)
        cr.set_trunc_nums(trunc)
        if kind == 'flux_transmon':
            cr.set_charge_offset(1, 0.17)
        names = list(els)

        def grads():
            g = np.array([0.0 if els[n].grad is None else float(els[n].grad) for n in names])
            cr.zero_grad()
            return g
        ef, _ = cr.diag(k)
        rec[kind + '/names'] = np.array(names)
        rec[kind + '/efreqs'] = ef.detach().numpy()
        (ef[1] - ef[0]).backward()
        rec[kind + '/grad_f10'] = grads()
        for d in decs:
            cr.diag(k)
            r = cr.dec_rate(d, (0, 1))
            r.backward()
            rec[kind + '/dec_' + d] = float(r)
            rec[kind + '/grad_dec_' + d] = grads()
    for idx in range(3):
        kind = 'discovery%d' % idx
        cr, els, trunc, k = grad_circuits(sq, kind)
        cr.set_trunc_nums(trunc)
        names = list(els)
        rec[kind + '/names'] = np.array(names)
        G = []
        for m in range(k):
            ef, _ = cr.diag(k)
            if m == 0:
                rec[kind + '/efreqs'] = ef.detach().numpy()
            ef[m].backward()
            G.append([0.0 if els[n].grad is None else float(els[n].grad) for n in names])
            cr.zero_grad()
        rec[kind + '/grad_efreqs'] = np.array(G)           # [k][P]
    set_optim_mode(False)
    np.savez_compressed(os.path.join(HERE, 'gradients.npz'), **rec)


PARTS = dict(grad=part_grad, zeropi=part_zeropi, fluxonium=part_fluxonium, transmon=part_transmon,
             coupled=part_coupled, trt=part_trt, zeropi_big=part_zeropi_big)

if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else None)
