"""Kernel-level parity on the GPU: every C entry point against a torch/numpy restatement
(the same contracts tests/emul_kernels.py states), through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from sqcircuit_b200.kernels import BlockView, Kernels, make_space  # noqa: E402
from emul_kernels import EmulKernels  # noqa: E402


@pytest.fixture(scope='module')
def kern():
    assert torch.cuda.is_available()
    return Kernels('cuda:0')


def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(shape, dtype=torch.float64, generator=g),
                         torch.randn(shape, dtype=torch.float64, generator=g))


def relerr(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))


@pytest.mark.parametrize('B,na,nb,N', [(3, 48, 12, 10100), (5, 7, 3, 625), (2, 64, 16, 1000),
                                       (300, 24, 12, 777), (1, 12, 12, 199), (4, 33, 9, 4096),
                                       (7, 2, 2, 10100), (5, 1, 1, 333), (3, 4, 3, 1000), (6, 2, 1, 64), (2, 3, 4, 5)])
@pytest.mark.parametrize('impl', [0, 1])
def test_gram(kern, B, na, nb, N, impl):
    A = rnd((B, na, N), 1)
    Bm = rnd((B, nb, N), 2)
    ca = torch.randint(0, na + 1, (B,), dtype=torch.int32)
    ca[0] = na
    cb = torch.randint(0, nb + 1, (B,), dtype=torch.int32)
    cb[0] = nb
    ref = torch.zeros((B, na, 16), dtype=torch.complex128)
    EmulKernels().gram(BlockView(A, cnt=ca), BlockView(Bm, cnt=cb), ref)
    out = torch.full((B, na, 16), 7.0, dtype=torch.complex128, device='cuda')
    kern.gram(BlockView(A.cuda(), cnt=ca.cuda()), BlockView(Bm.cuda(), cnt=cb.cuda()), out, impl)
    assert relerr(out.cpu()[:, :na, :nb], ref[:, :na, :nb]) < 1e-13


def test_gram_offsets(kern):
    B, mm, N = 4, 40, 999
    V = rnd((B, mm, N), 3)
    off = torch.tensor([10, 20, 28, 5], dtype=torch.int32)
    cnt = torch.tensor([12, 3, 12, 0], dtype=torch.int32)
    msz = off.clone()
    ref = torch.zeros((B, mm, 16), dtype=torch.complex128)
    EmulKernels().gram(BlockView(V, cnt_fixed=mm, cnt=msz), BlockView(V, cnt_fixed=12, off=off, cnt=cnt), ref)
    Vd = V.cuda()
    out = torch.zeros((B, mm, 16), dtype=torch.complex128, device='cuda')
    kern.gram(BlockView(Vd, cnt_fixed=mm, cnt=msz.cuda()),
              BlockView(Vd, cnt_fixed=12, off=off.cuda(), cnt=cnt.cuda()), out)
    assert relerr(out.cpu(), ref) < 1e-13


@pytest.mark.parametrize('nb', [1, 3, 8, 10, 12, 16])
@pytest.mark.parametrize('op,trans', [(0, 0), (1, 1)])
@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('na,N', [(37, 2050), (48, 10100), (5, 131)])
def test_block_update(kern, nb, op, trans, impl, na, N):
    B = 3
    A = rnd((B, na, N), 4)
    Out = rnd((B, nb, N), 5)
    Cf = rnd((B, 48, 48), 6)
    ca = torch.tensor([na, na // 2, 0], dtype=torch.int32)
    cb = torch.tensor([nb, max(nb - 2, 1), nb], dtype=torch.int32)
    ref = Out.clone()
    EmulKernels().block_update(BlockView(A, cnt=ca), Cf, BlockView(ref, cnt=cb), op, trans)
    o = Out.cuda()
    kern.block_update(BlockView(A.cuda(), cnt=ca.cuda()), Cf.cuda(), BlockView(o, cnt=cb.cuda()), op, trans,
                      impl)
    assert relerr(o.cpu(), ref) < 1e-13


def test_block_rightmul(kern):
    B, N = 3, 1500
    T = rnd((B, 40, N), 7)
    off = torch.tensor([5, 28, 0], dtype=torch.int32)
    nin = torch.tensor([12, 7, 0], dtype=torch.int32)
    nout = torch.tensor([9, 7, 0], dtype=torch.int32)
    Cq = rnd((B, 16, 16), 8)
    ref = T.clone()
    EmulKernels().block_rightmul(BlockView(ref, cnt_fixed=12, off=off, cnt=nin), Cq, nout)
    t = T.cuda()
    kern.block_rightmul(BlockView(t, cnt_fixed=12, off=off.cuda(), cnt=nin.cuda()), Cq.cuda(), nout.cuda())
    assert relerr(t.cpu(), ref) < 1e-13


@pytest.mark.parametrize('ld,sizes', [(48, [48, 31, 12, 1, 0, 47]), (16, [16, 5]), (100, [100, 61]),
                                      (112, [112]), (64, [64, 63])])
def test_heev(kern, ld, sizes):
    B = len(sizes)
    A = rnd((B, ld, ld), 9)
    A = A + A.conj().transpose(1, 2)
    # a spectrum with clusters and a wide range, like a projected Hamiltonian
    A[0] = A[0] * 1e10
    n = torch.tensor(sizes, dtype=torch.int32)
    w = torch.full((B, ld), -1.0, dtype=torch.float64, device='cuda')
    Zt = torch.zeros((B, ld, ld), dtype=torch.complex128, device='cuda')
    kern.heev(A.cuda(), n.cuda(), w, Zt)
    w, Zt = w.cpu(), Zt.cpu()
    for b, nb in enumerate(sizes):
        if nb == 0:
            assert float(w[b, 0]) == -1.0
            continue
        a = A[b, :nb, :nb].numpy()
        wr = np.linalg.eigvalsh(a)
        scale = np.abs(wr).max()
        assert np.abs(w[b, :nb].numpy() - wr).max() < 1e-13 * scale
        Z = Zt[b, :nb, :nb].numpy().T
        assert np.abs(Z.conj().T @ Z - np.eye(nb)).max() < 1e-13
        assert np.abs(a @ Z - Z * w[b, :nb].numpy()[None, :]).max() < 2e-13 * scale * np.sqrt(nb)


@pytest.mark.parametrize('n', [2, 7, 50, 100, 111, 112])
def test_syev(kern, n):
    """Batched real symmetric Jacobi eigensolver (sqc_syev) against torch.linalg.eigh in float64: eigenvalues,
    residuals, orthonormality; includes a matrix with degenerate eigenvalues and a diagonal one."""
    B = 5
    gen = torch.Generator().manual_seed(100 + n)
    A = torch.randn((B, n, n), dtype=torch.float64, generator=gen)
    A = A + A.transpose(1, 2)
    A[1] = torch.diag(torch.arange(n, dtype=torch.float64) - 3.0)                      # already diagonal
    q, _ = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, generator=gen))
    A[2] = q @ torch.diag(torch.tensor([1.0, 1.0] + [float(i // 2) for i in range(2, n)], dtype=torch.float64)[:n]) @ q.T
    A[3] = A[3] * 1e9                                                                  # circuit-like scale (rad/s)
    Ad = A.cuda().contiguous()
    w = torch.empty((B, n), dtype=torch.float64, device='cuda')
    Zt = torch.empty((B, n, n), dtype=torch.float64, device='cuda')
    kern.syev(Ad, w, Zt)
    ref = torch.linalg.eigvalsh(0.5 * (A + A.transpose(1, 2)))
    for b in range(B):
        scale = float(ref[b].abs().max())
        assert float((w[b].cpu() - ref[b]).abs().max()) < 1e-12 * scale
        Z = Zt[b].cpu()
        assert float((Z @ Z.T - torch.eye(n, dtype=torch.float64)).abs().max()) < 1e-12
        As = 0.5 * (A[b] + A[b].T)
        assert float((Z @ As - w[b].cpu()[:, None] * Z).abs().max()) < 1e-11 * scale


@pytest.mark.parametrize('m', [2, 9, 25, 100, 101, 112])
def test_xbasis(kern, m):
    V, lam = kern.xbasis(m)
    V, lam = V.cpu().numpy(), lam.cpu().numpy()
    x = np.diag(np.sqrt(np.arange(1, m)), 1)
    x = x + x.T
    assert np.all(np.diff(lam) > 0)
    assert np.abs(np.linalg.eigvalsh(x) - lam).max() < 1e-13 * np.sqrt(m) * 4
    assert np.abs(V.T @ V - np.eye(m)).max() < 1e-13
    assert np.abs(x @ V - V * lam[None, :]).max() < 1e-12
    # displacement operator D(i s) == expm(i s (a + a^dag))  (what qt.displace computes)
    import scipy.linalg
    s = 0.37
    D = (V * np.exp(1j * s * lam)[None, :]) @ V.T
    assert np.abs(D - scipy.linalg.expm(1j * s * x)).max() < 1e-12


def test_svqb(kern):
    B, N, p = 4, 800, 12
    T = rnd((B, p, N), 11)
    T[1, 5] = T[1, 2] * (1 + 1e-9) + 1e-15 * T[1, 3]     # (numerically) dependent direction
    T[2, 7] = 0                                           # zero vector
    nact = torch.tensor([12, 12, 12, 0], dtype=torch.int32)
    msz = torch.tensor([12, 24, 36, 12], dtype=torch.int32)
    Td = T.cuda()
    S = torch.zeros((B, 16, 16), dtype=torch.complex128, device='cuda')
    nin = nact.clone().cuda()
    na = nact.clone().cuda()
    kern.gram(BlockView(Td, cnt=nin), BlockView(Td, cnt=nin), S)
    Cq = torch.zeros((B, 16, 16), dtype=torch.complex128, device='cuda')
    mtot = torch.zeros(B, dtype=torch.int32, device='cuda')
    kern.svqb_coeffs(S, na, 1e-12, Cq, msz.cuda(), mtot)
    kern.block_rightmul(BlockView(Td, cnt=nin), Cq, na)
    na = na.cpu()
    assert list(na) == [12, 11, 11, 0]
    assert list(mtot.cpu()) == [24, 35, 47, 12]
    for b in range(3):
        q = Td[b, :int(na[b])].cpu().numpy()
        assert np.abs(q.conj() @ q.T - np.eye(int(na[b]))).max() < 1e-6   # first SVQB pass
        assert float(Td[b, int(na[b]):p].abs().max()) == 0.0 if int(na[b]) < p else True


def _space_cases():
    return [([25, 25], [1, 0]), ([100, 101], [1, 0]), ([6, 11, 11], [1, 0, 0]), ([100], [1]),
            ([1, 9, 23], [1, 1, 1]), ([61], [0]), ([7, 5, 9], [0, 1, 0])]


@pytest.mark.parametrize('dims,harm', _space_cases())
@pytest.mark.parametrize('impl', [0, 1])
def test_mode_transform(kern, dims, harm, impl):
    N = int(np.prod(dims))
    B, nv = 3, 5
    sp = make_space(dims, harm)
    X = rnd((B, nv, N), 12)
    cnt = torch.tensor([5, 2, 0], dtype=torch.int32)
    ed = torch.randn((B, N), dtype=torch.float64)
    for mode, d in enumerate(dims):
        if not harm[mode] or d == 1:
            continue
        mat = torch.randn((d, d), dtype=torch.float64)
        for transpose in (0, 1):
            for use_ep in (False, True):
                ref = torch.zeros_like(X)
                EmulKernels().mode_transform(sp, mode, mat, transpose, BlockView(X, cnt=cnt),
                                             BlockView(ref, cnt=cnt), ed if use_ep else None,
                                             BlockView(X, cnt=cnt) if use_ep else None)
                Xd = X.cuda()
                out = torch.zeros_like(Xd)
                kern.mode_transform(sp, mode, mat.cuda(), transpose, BlockView(Xd, cnt=cnt.cuda()),
                                    BlockView(out, cnt=cnt.cuda()), ed.cuda() if use_ep else None,
                                    BlockView(Xd, cnt=cnt.cuda()) if use_ep else None, impl=impl)
                assert relerr(out.cpu(), ref) < 1e-13, (mode, transpose, use_ep)
                if not use_ep:  # in place
                    kern.mode_transform(sp, mode, mat.cuda(), transpose, BlockView(Xd, cnt=cnt.cuda()),
                                        BlockView(Xd, cnt=cnt.cuda()), impl=impl)
                    got = Xd.cpu()
                    assert relerr(got[0], ref[0]) < 1e-13 and relerr(got[1, :2], ref[1, :2]) < 1e-13
                    assert relerr(got[1, 2:], X[1, 2:]) == 0 and relerr(got[2], X[2]) == 0


def test_small_elementwise_kernels(kern):
    B, p, N = 3, 6, 1000
    X, HX = rnd((B, p, N), 13), rnd((B, p, N), 14)
    theta = torch.randn((B, 48), dtype=torch.float64)
    E = EmulKernels()
    ref = torch.zeros((B, p), dtype=torch.float64)
    E.residual_norms(BlockView(X), BlockView(HX), theta, ref)
    out = torch.zeros((B, p), dtype=torch.float64, device='cuda')
    kern.residual_norms(BlockView(X.cuda()), BlockView(HX.cuda()), theta.cuda(), out)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-13
    # preconditioned residual into the tail of a basis
    V = rnd((B, 30, N), 15)
    msz = torch.tensor([6, 12, 20], dtype=torch.int32)
    nact = torch.tensor([6, 3, 0], dtype=torch.int32)
    act = torch.tensor([[0, 1, 2, 3, 4, 5], [5, 0, 2, 0, 0, 0], [0] * 6], dtype=torch.int32)
    d = torch.randn((B, N), dtype=torch.float64) * 10
    fl = torch.tensor([0.5, 0.1, 0.1], dtype=torch.float64)
    refV = V.clone()
    E.precond_residual(BlockView(X), BlockView(HX), BlockView(refV, cnt_fixed=p, off=msz, cnt=nact), theta, act, d, fl)
    Vd = V.cuda()
    kern.precond_residual(BlockView(X.cuda()), BlockView(HX.cuda()),
                          BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()), theta.cuda(),
                          act.cuda(), d.cuda(), fl.cuda())
    assert relerr(Vd.cpu(), refV) < 1e-13
    # restart copy
    W = rnd((B, 30, N), 16)
    rs = torch.tensor([1, 0, 1], dtype=torch.int32)
    refV2, refW2 = V.clone(), W.clone()
    E.restart_copy(BlockView(X), BlockView(HX), BlockView(refV2, cnt_fixed=p), BlockView(refW2, cnt_fixed=p), rs)
    Vd, Wd = V.cuda(), W.cuda()
    kern.restart_copy(BlockView(X.cuda()), BlockView(HX.cuda()), BlockView(Vd, cnt_fixed=p),
                      BlockView(Wd, cnt_fixed=p), rs.cuda())
    assert relerr(Vd.cpu(), refV2) == 0 and relerr(Wd.cpu(), refW2) == 0
    # diag mul
    Y = rnd((B, p, N), 17)
    refY = Y.clone()
    E.diag_mul(BlockView(X), BlockView(refY), d, True)
    Yd = Y.cuda()
    kern.diag_mul(BlockView(X.cuda()), BlockView(Yd), d.cuda(), True)
    assert relerr(Yd.cpu(), refY) < 1e-14


@pytest.mark.parametrize('dims', [[100, 101], [25, 25], [30, 19], [7, 3], [2, 61], [101, 40]])
def test_hx_fused_matches_unfused(kern, dims):
    """Fused H.X kernel against the transform / potential / transform sequence."""
    import ctypes as C
    from sqcircuit_b200._lib import Potential
    m, inner = dims
    N = m * inner
    B, nv, T = 3, 4, 2
    sp = make_space(dims, [1, 0])
    V, lam = kern.xbasis(m)
    lamcat = torch.zeros(m + inner, dtype=torch.float64, device='cuda')
    lamcat[:m] = lam
    g = torch.Generator().manual_seed(5)
    s = torch.tensor([[[0.61, 0.0], [-0.61, 0.0]]], dtype=torch.float64, device='cuda')
    etab = torch.empty((1, T, m + inner), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lamcat, s, etab)
    a = rnd((B, T), 21).cuda()
    gl = torch.randn((B, 3), dtype=torch.float64, generator=g).cuda()
    fd = torch.randn((B, N), dtype=torch.float64, generator=g).cuda()
    X = rnd((B, nv, N), 22).cuda()
    cnt = torch.tensor([4, 1, 0], dtype=torch.int32, device='cuda')

    def pot_struct():
        pot = Potential()
        pot.n_terms = T
        pot.shift[0][1] = 1
        pot.shift[1][1] = -1
        pot.lam, pot.etab, pot.etab_stride_b = lamcat.data_ptr(), etab.data_ptr(), 0
        pot.g, pot.a, pot.diag, pot.diag_stride_b = gl.data_ptr(), a.data_ptr(), None, 0
        return pot
    for use_fd in (True, False):
        ref = torch.zeros_like(X)
        s1, s2 = torch.zeros_like(X), torch.zeros_like(X)
        kern.mode_transform(sp, 0, V, 1, BlockView(X, cnt=cnt), BlockView(s1, cnt=cnt), impl=1)
        kern.potential_apply(sp, pot_struct(), BlockView(s1, cnt=cnt), BlockView(s2, cnt=cnt))
        if use_fd:
            kern.mode_transform(sp, 0, V, 0, BlockView(s2, cnt=cnt), BlockView(ref, cnt=cnt), ediag=fd,
                                EX=BlockView(X, cnt=cnt), impl=1)
        else:
            kern.mode_transform(sp, 0, V, 0, BlockView(s2, cnt=cnt), BlockView(ref, cnt=cnt), impl=1)
        out = torch.zeros_like(X)
        he, ho = (m + 1) // 2, m // 2
        ok = kern.hx_fused(sp, V[0::2, :he].contiguous(), V[1::2, :ho].contiguous(), pot_struct(),
                           fd if use_fd else None, BlockView(X, cnt=cnt), BlockView(out, cnt=cnt))
        assert ok
        assert relerr(out.cpu(), ref.cpu()) < 2e-13, (dims, use_fd)


def test_lowblock_preconditioner_kernels(kern):
    """assemble / inverse / apply of the two-level preconditioner against the numpy contracts."""
    from sqcircuit_b200._lib import Potential
    m, inner, ns, B, T = 25, 25, 96, 3, 2
    N = m * inner
    sp = make_space([m, inner], [1, 0])
    V, lam = kern.xbasis(m)
    lamcat = torch.zeros(m + inner, dtype=torch.float64, device='cuda')
    lamcat[:m] = lam
    s = torch.tensor([[[0.61, 0.0], [-0.61, 0.0]]], dtype=torch.float64, device='cuda')
    etab = torch.empty((1, T, m + inner), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lamcat, s, etab)
    Dt = torch.einsum('ni,ti,ki->tnk', V.to(torch.complex128), etab[0, :, :m], V.to(torch.complex128)).contiguous()
    a = rnd((B, T), 31).cuda()
    gl = torch.randn((B, 3), dtype=torch.float64, generator=torch.Generator().manual_seed(3)).cuda()
    d = (torch.rand((1, N), dtype=torch.float64, generator=torch.Generator().manual_seed(4)) * 50 + 20).cuda()
    sigma = torch.tensor([-5.0, -6.0, -7.0], dtype=torch.float64, device='cuda')
    S = torch.topk(d, ns, dim=1, largest=False).indices.to(torch.int32).contiguous()

    def pot_struct(dev_a, dev_g, dev_e):
        pot = Potential()
        pot.n_terms = T
        pot.shift[0][1] = 1
        pot.shift[1][1] = -1
        pot.lam, pot.etab, pot.etab_stride_b = lamcat.data_ptr(), dev_e.data_ptr(), 0
        pot.g, pot.a, pot.diag, pot.diag_stride_b = dev_g.data_ptr(), dev_a.data_ptr(), None, 0
        pot._keep = (dev_a, dev_g, dev_e, None)
        return pot
    out = torch.empty((B, ns, ns), dtype=torch.complex64, device='cuda')
    kern.lowblock_assemble(sp, S, Dt, pot_struct(a, gl, etab), d, sigma, 1.0, out)
    ref = torch.empty((B, ns, ns), dtype=torch.complex64)
    E = EmulKernels()
    E.lowblock_assemble(sp, S.cpu(), Dt.cpu(), pot_struct(a.cpu(), gl.cpu(), etab.cpu()), d.cpu(), sigma.cpu(), 1.0, ref)
    assert relerr(out.cpu().to(torch.complex128), ref.to(torch.complex128)) < 1e-6
    # Hermitian, diagonally dominant here -> positive definite; inverse in shared memory
    inv = out.clone()
    kern.hpd_inverse_c64(inv)
    prod = inv.to(torch.complex128) @ out.to(torch.complex128)
    assert float((prod - torch.eye(ns, device='cuda')).abs().max()) < 5e-4
    # apply
    p = 6
    X, HX = rnd((B, p, N), 32).cuda(), rnd((B, p, N), 33).cuda()
    Vb = rnd((B, 20, N), 34)
    theta = torch.randn((B, 48), dtype=torch.float64)
    msz = torch.tensor([6, 10, 14], dtype=torch.int32)
    nact = torch.tensor([6, 2, 0], dtype=torch.int32)
    act = torch.tensor([[0, 1, 2, 3, 4, 5], [4, 1, 0, 0, 0, 0], [0] * 6], dtype=torch.int32)
    refV = Vb.clone()
    E.lowblock_apply(BlockView(X.cpu()), BlockView(HX.cpu()), BlockView(refV, cnt_fixed=p, off=msz, cnt=nact),
                     theta, act, S.cpu(), inv.cpu(), 2.0)
    Vd = Vb.cuda()
    kern.lowblock_apply(BlockView(X), BlockView(HX), BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                        theta.cuda(), act.cuda(), S, inv, 2.0)
    assert relerr(Vd.cpu(), refV) < 1e-12
    # grouped: sweep points b share the inverse of slot b // group
    refV = Vb.clone()
    E.lowblock_apply(BlockView(X.cpu()), BlockView(HX.cpu()), BlockView(refV, cnt_fixed=p, off=msz, cnt=nact),
                     theta, act, S.cpu(), inv.cpu()[:2], 2.0, group=2)
    Vd = Vb.cuda()
    kern.lowblock_apply(BlockView(X), BlockView(HX), BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                        theta.cuda(), act.cuda(), S, inv[:2].contiguous(), 2.0, group=2)
    assert relerr(Vd.cpu(), refV) < 1e-12


def test_gram_pair_and_gen_rr(kern):
    """Paired Gram pass and the generalised Rayleigh-Ritz kernel against the numpy contracts."""
    E = EmulKernels()
    B, mm, N, p = 5, 48, 3001, 12
    V = rnd((B, mm, N), 41)
    W = rnd((B, mm, N), 42)
    msz = torch.tensor([12, 24, 36, 30, 0], dtype=torch.int32)
    nact = torch.tensor([12, 7, 12, 1, 0], dtype=torch.int32)
    mtot = msz + nact
    r1, r2 = torch.zeros((B, mm, 16), dtype=torch.complex128), torch.zeros((B, mm, 16), dtype=torch.complex128)
    E.gram_pair(BlockView(V, cnt_fixed=mm, cnt=mtot), BlockView(V, cnt_fixed=p, off=msz, cnt=nact),
                BlockView(W, cnt_fixed=p, off=msz, cnt=nact), r1, r2)
    Vd, Wd = V.cuda(), W.cuda()
    o1 = torch.full((B, mm, 16), 3.0, dtype=torch.complex128, device='cuda')
    o2 = torch.full((B, mm, 16), 3.0, dtype=torch.complex128, device='cuda')
    kern.gram_pair(BlockView(Vd, cnt_fixed=mm, cnt=mtot.cuda()),
                   BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                   BlockView(Wd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()), o1, o2)
    assert relerr(o1.cpu()[:, :, :p], r1[:, :, :p]) < 1e-13 and relerr(o2.cpu()[:, :, :p], r2[:, :, :p]) < 1e-13
    # generalised RR on a non-orthogonal basis with one exactly dependent vector
    n = 40
    Q = rnd((B, n, 200), 43)
    Q[1, 7] = Q[1, 3] * (0.5 + 0.2j) + Q[1, 5]          # dependent -> must be removed
    Hs = rnd((B, 200, 200), 44)
    Hs = Hs + Hs.conj().transpose(1, 2)
    GB = Q.conj() @ Q.transpose(1, 2)
    GA = Q.conj() @ Hs @ Q.transpose(1, 2)
    GBp = torch.zeros((B, mm, mm), dtype=torch.complex128)
    GAp = torch.zeros((B, mm, mm), dtype=torch.complex128)
    GBp[:, :n, :n], GAp[:, :n, :n] = GB, GA
    nn = torch.tensor([n, n, 12, 1, 0], dtype=torch.int32)
    w = torch.full((B, mm), -7.0, dtype=torch.float64, device='cuda')
    Zt = torch.zeros((B, mm, mm), dtype=torch.complex128, device='cuda')
    kern.gen_rr(GAp.cuda(), GBp.cuda(), nn.cuda(), 1e-8, w, Zt)
    w, Zt = w.cpu(), Zt.cpu()
    import scipy.linalg
    for b in range(4):
        nb = int(nn[b])
        a, g = GA[b, :nb, :nb].numpy(), GB[b, :nb, :nb].numpy()
        if b == 1:
            keep = [i for i in range(nb) if i != 7]
            wr = scipy.linalg.eigh(a[np.ix_(keep, keep)], g[np.ix_(keep, keep)], eigvals_only=True)
        else:
            wr = scipy.linalg.eigh(a, g, eigvals_only=True)
        k = len(wr)
        assert np.abs(w[b, :k].numpy() - wr).max() < 1e-10 * np.abs(wr).max()
        c = Zt[b, :k, :nb].numpy().T                   # columns = coefficient vectors
        assert np.abs(c.conj().T @ g @ c - np.eye(k)).max() < 1e-9
        assert np.abs(a @ c - g @ c * w[b, :k].numpy()[None, :]).max() < 1e-9 * np.abs(wr).max() * np.abs(c).max()
    assert float(w[4, 0]) == -7.0


def test_launch_counter(kern):
    n0 = kern.launch_count()
    kern.xbasis(8)
    assert kern.launch_count() == n0 + 1


@pytest.mark.parametrize('planned', [False, True])
def test_davidson_decide_matches_emulation(kern, planned):
    """Bookkeeping kernel against the emulation, legacy and copy-free (planned) restart modes,
    over a few chained calls so that the planned flags of one call drive the next."""
    B, p, k, mmax, ld = 40, 6, 4, 20, 20
    gen = torch.Generator().manual_seed(5)

    def fresh(dev):
        t = dict(msz=torch.randint(p, mmax + 1, (B,), generator=torch.Generator().manual_seed(1)).to(torch.int32),
                 nact=torch.zeros(B, dtype=torch.int32), act=torch.zeros((B, p), dtype=torch.int32),
                 restart=torch.zeros(B, dtype=torch.int32), done=torch.zeros(B, dtype=torch.int32),
                 n_not_done=torch.zeros(1, dtype=torch.int32),
                 theta=torch.sort(torch.randn((B, ld), dtype=torch.float64, generator=torch.Generator().manual_seed(2)), dim=1).values,
                 rn2=torch.zeros((B, p), dtype=torch.float64), resmax=torch.zeros(B, dtype=torch.float64),
                 den_floor=torch.zeros(B, dtype=torch.float64),
                 G=torch.zeros((B, ld, ld), dtype=torch.complex128), GB=torch.zeros((B, ld, ld), dtype=torch.complex128),
                 mtot=torch.zeros(B, dtype=torch.int32))
        if planned:
            t['xoff'] = torch.full((B,), mmax, dtype=torch.int32)
            t['xlast'] = torch.full((B,), mmax, dtype=torch.int32)
        return {kk: v.to(dev) for kk, v in t.items()}
    tc, tg = fresh('cpu'), fresh('cuda')
    # a few points carry non-finite wanted / guard Ritz values (dependent vectors dropped by the generalised
    # Rayleigh-Ritz): they must stay unfinished and give finite preconditioner floors
    for t in (tc, tg):
        t['theta'][3, k - 1:] = float('inf')
        t['theta'][7, p - 1:] = float('inf')
        t['theta'][11, 0:] = float('inf')
    E = EmulKernels()
    extra = dict(xrow=mmax) if planned else {}
    sc = E.davidson_state(p=p, k=k, mmax=mmax, ldg=ld, tol_rel=1e-3, tol_abs_floor=1e-300, **tc, **extra)
    sg = kern.davidson_state(p=p, k=k, mmax=mmax, ldg=ld, tol_rel=1e-3, tol_abs_floor=1e-300, **tg, **extra)
    for it in range(4):
        rn = (10.0 ** (-torch.rand((B, p), dtype=torch.float64, generator=gen) * (3 + 2 * it))) ** 2
        tc['rn2'].copy_(rn)
        tg['rn2'].copy_(rn.cuda())
        E.davidson_decide(sc, B)
        kern.davidson_decide(sg, B)
        assert not bool(tg['done'][3]) and not bool(tg['done'][11]) and bool(torch.isfinite(tg['den_floor']).all())
        for name in tc:
            a, b_ = tc[name], tg[name].cpu()
            if name == 'theta':
                assert torch.equal(a, b_)
            elif name == 'act':       # entries past nact are unspecified
                for b in range(B):
                    n = int(tc['nact'][b])
                    assert torch.equal(a[b, :n], b_[b, :n])
            elif a.dtype.is_floating_point or a.dtype.is_complex:
                fin = torch.isfinite(a.real) if a.dtype.is_complex else torch.isfinite(a)
                assert torch.equal(torch.isfinite(b_.real if a.dtype.is_complex else b_), fin), name
                zero = torch.zeros((), dtype=a.dtype)
                assert relerr(torch.where(fin, b_, zero), torch.where(fin, a, zero)) < 1e-14, name
            else:
                assert torch.equal(a, b_), name
        # emulate the insert: the basis grows by nact
        tc['msz'] += tc['nact']
        tg['msz'] += tg['nact']
        assert int(tc['msz'].max()) <= mmax


@pytest.mark.parametrize('B,mm,p,N', [(3, 40, 12, 10100), (5, 20, 6, 777), (2, 64, 16, 1000), (4, 12, 4, 130)])
def test_ritz_update(kern, B, mm, p, N):
    """Fused X = C V, HX = C W, ||HX - theta X||^2 against the emulation, with per-point basis
    sizes, an inactive point, and the Ritz block aliasing the leading basis rows (planned restart)."""
    V, W = rnd((B, mm + p, N), 41), rnd((B, mm + p, N), 42)
    Cf = rnd((B, mm, mm), 43)
    theta = torch.randn((B, mm), dtype=torch.float64, generator=torch.Generator().manual_seed(44))
    msz = torch.randint(p, mm + 1, (B,), generator=torch.Generator().manual_seed(45)).to(torch.int32)
    msz[0] = mm
    xcnt = torch.full((B,), p, dtype=torch.int32)
    xcnt[-1] = 0                                     # inactive sweep point
    xoff = torch.full((B,), mm, dtype=torch.int32)
    xoff[0] = 0                                      # in place over the first p basis vectors
    rv, rw = V.clone(), W.clone()
    rn_ref = torch.full((B, p), 3.0, dtype=torch.float64)
    EmulKernels().ritz_update(BlockView(rv, cnt_fixed=mm, cnt=msz), BlockView(rw, cnt_fixed=mm, cnt=msz), Cf,
                              BlockView(rv, cnt_fixed=p, off=xoff, cnt=xcnt),
                              BlockView(rw, cnt_fixed=p, off=xoff, cnt=xcnt), theta, rn_ref)
    Vd, Wd = V.cuda(), W.cuda()
    rn = torch.full((B, p), 3.0, dtype=torch.float64, device='cuda')
    mszd, xcd, xod = msz.cuda(), xcnt.cuda(), xoff.cuda()
    kern.ritz_update(BlockView(Vd, cnt_fixed=mm, cnt=mszd), BlockView(Wd, cnt_fixed=mm, cnt=mszd), Cf.cuda(),
                     BlockView(Vd, cnt_fixed=p, off=xod, cnt=xcd), BlockView(Wd, cnt_fixed=p, off=xod, cnt=xcd),
                     theta.cuda(), rn)
    assert relerr(Vd.cpu(), rv) < 1e-13 and relerr(Wd.cpu(), rw) < 1e-13
    assert float((rn.cpu() - rn_ref).abs().max() / rn_ref.abs().max()) < 1e-12
    assert torch.equal(rn.cpu()[-1], torch.full((p,), 3.0, dtype=torch.float64))


def test_lowband_factor_and_apply(kern):
    """Banded two-level preconditioner: block-LDL^H factor and substitution against the emulation and
    against a dense solve of the same block-tridiagonal Hermitian positive definite matrices."""
    rng = np.random.default_rng(7)
    sizes = [[3, 7, 12, 17, 12, 6, 3], [5, 5, 20, 20, 10], [60]]
    B, ns, N, p = 5, 60, 500, 6
    nslot, group = 3, 2
    off = np.zeros((nslot, 8), dtype=np.int32)
    nblk = np.zeros(nslot, dtype=np.int32)
    A = np.zeros((nslot, ns, ns), dtype=complex)
    for sl, sz in enumerate(sizes):
        o = np.concatenate(([0], np.cumsum(sz)))
        off[sl, :len(o)] = o
        nblk[sl] = len(sz)
        for q in range(len(sz)):
            a, b_ = o[q], o[q + 1]
            d = rng.standard_normal((sz[q], sz[q])) + 1j * rng.standard_normal((sz[q], sz[q]))
            A[sl, a:b_, a:b_] = d @ d.conj().T / sz[q] + 4 * np.eye(sz[q])
            if q + 1 < len(sz):
                c = o[q + 2]
                e = 0.5 * (rng.standard_normal((sz[q], sz[q + 1])) + 1j * rng.standard_normal((sz[q], sz[q + 1])))
                A[sl, a:b_, b_:c] = e
                A[sl, b_:c, a:b_] = e.conj().T
    A64 = torch.from_numpy(A.astype(np.complex64))
    offt, nbt = torch.from_numpy(off), torch.from_numpy(nblk)
    E = EmulKernels()
    fref = A64.clone()
    E.lowband_factor(fref, offt, nbt, 60)
    fac = A64.clone().cuda()
    kern.lowband_factor(fac, offt.cuda(), nbt.cuda(), 60)
    assert relerr(fac.cpu().to(torch.complex128), fref.to(torch.complex128)) < 2e-5
    S = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(s))[:ns] for s in range(nslot)]).to(torch.int32)
    X, HX = rnd((B, p, N), 51), rnd((B, p, N), 52)
    Vb = rnd((B, 20, N), 53)
    theta = torch.randn((B, 48), dtype=torch.float64, generator=torch.Generator().manual_seed(54))
    msz = torch.tensor([6, 10, 14, 3, 8], dtype=torch.int32)
    nact = torch.tensor([6, 2, 0, 4, 1], dtype=torch.int32)
    act = torch.tensor([[0, 1, 2, 3, 4, 5], [4, 1, 0, 0, 0, 0], [0] * 6, [5, 3, 2, 0, 0, 0], [2, 0, 0, 0, 0, 0]], dtype=torch.int32)
    refV = Vb.clone()
    E.lowband_apply(BlockView(X), BlockView(HX), BlockView(refV, cnt_fixed=p, off=msz, cnt=nact), theta, act, S,
                    fref, offt, nbt, 60, 2.0, group=group)
    Vd = Vb.cuda()
    kern.lowband_apply(BlockView(X.cuda()), BlockView(HX.cuda()), BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                       theta.cuda(), act.cuda(), S.cuda(), fac, offt.cuda(), nbt.cuda(), 60, 2.0, group=group)
    assert relerr(Vd.cpu(), refV) < 2e-5
    # address-ordered gather / scatter gives the same result
    Vd2 = Vb.cuda()
    kern.lowband_apply(BlockView(X.cuda()), BlockView(HX.cuda()), BlockView(Vd2, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                       theta.cuda(), act.cuda(), S.cuda(), fac, offt.cuda(), nbt.cuda(), 60, 2.0, group=group,
                       perm=torch.argsort(S, dim=1).to(torch.int32).cuda())
    assert torch.equal(Vd2, Vd)
    # the substitution really solves the block-tridiagonal system
    b, jj = 0, 1
    s = S[0].long().numpy()
    r = (HX[b, jj].numpy()[s] - float(theta[b, jj]) * X[b, jj].numpy()[s]) / 2.0
    sol = np.linalg.solve(A[0], r)
    got = Vd.cpu()[b, int(msz[b]) + jj].numpy()[s]
    assert np.abs(got - sol).max() / np.abs(sol).max() < 1e-4


def _real_case(kern, m, inner, B, T=2, seed=40):
    """Random harmonic x charge operator tables with a charge-symmetric LC diagonal (zero gate charge)."""
    from sqcircuit_b200._lib import Potential
    N = m * inner
    nc = (inner - 1) // 2
    sp = make_space([m, inner], [1, 0])
    V, lam = kern.xbasis(m)
    lamcat = torch.zeros(m + inner, dtype=torch.float64, device='cuda')
    lamcat[:m] = lam
    s = torch.tensor([[[0.61, 0.0], [-0.37, 0.0]]], dtype=torch.float64, device='cuda')
    etab = torch.empty((1, T, m + inner), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lamcat, s, etab)
    g = torch.Generator().manual_seed(seed)
    a = rnd((B, T), seed + 1).cuda()
    gl = torch.randn((B, 3), dtype=torch.float64, generator=g).cuda()
    half = torch.randn((1, m, nc + 1), dtype=torch.float64, generator=g)
    fdc = torch.cat([half.flip(2)[:, :, :nc], half], dim=2).reshape(1, N).contiguous().cuda()   # d(i, -n) = d(i, n)
    cmap = torch.as_tensor(np.concatenate(([nc], nc + np.arange(1, nc + 1), nc + np.arange(1, nc + 1))))
    fdr = fdc.reshape(1, m, inner)[:, :, cmap.cuda()].reshape(1, N)
    if N % 2:
        fdr = torch.cat([fdr, torch.full((1, 1), 1e300, dtype=torch.float64, device='cuda')], dim=1)
    fdr = fdr.contiguous()

    def pot_struct(shifts=(1, -1)):
        pot = Potential()
        pot.n_terms = T
        for t in range(T):
            pot.shift[t][1] = shifts[t]
        pot.lam, pot.etab, pot.etab_stride_b = lamcat.data_ptr(), etab.data_ptr(), 0
        pot.g, pot.a, pot.diag, pot.diag_stride_b = gl.data_ptr(), a.data_ptr(), None, 0
        pot._keep = (a, gl, etab)
        return pot
    return sp, V, etab, a, gl, fdc, fdr, pot_struct


@pytest.mark.parametrize('dims', [[100, 51], [25, 13], [30, 19], [7, 3], [101, 41], [2, 61], [112, 127], [33, 1],
                                  [100, 101]])
@pytest.mark.parametrize('shifts', [(1, -1), (1, 0)])
def test_hx_real_matches_complex(kern, dims, shifts):
    """Real-representation H.X (TMA staged) against the complex fused kernel on time-reversal
    symmetric vectors, through the conversion kernels."""
    m, inner = dims
    N, Nc = m * inner, (m * inner + 1) // 2
    nc_ = (inner - 1) // 2
    B, nv = 3, 5
    sp, V, etab, a, gl, fdc, fdr, pot_struct = _real_case(kern, m, inner, B)
    cnt = torch.tensor([5, 2, 0], dtype=torch.int32, device='cuda')
    Z = rnd((B, nv, N), 41).cuda()
    Xr = torch.zeros((B, nv, Nc), dtype=torch.complex128, device='cuda')
    kern.complex_to_real(sp, BlockView(Z, cnt=cnt), BlockView(Xr, cnt=cnt))
    Zs = torch.zeros_like(Z)
    kern.real_to_complex(sp, BlockView(Xr, cnt=cnt), BlockView(Zs, cnt=cnt))
    # the round trip is the projection on T-symmetric vectors: z(i, -n) = conj z(i, n)
    Zv = Zs.reshape(B, nv, m, inner)
    assert relerr(Zv.flip(3).conj().cpu(), Zv.cpu()) < 1e-15
    ref = 0.5 * (Z.reshape(B, nv, m, inner) + Z.reshape(B, nv, m, inner).flip(3).conj())
    live = (torch.arange(nv, device='cuda')[None, :] < cnt[:, None])[:, :, None, None]
    assert relerr((Zv * live).cpu(), (ref * live).cpu()) < 1e-15
    # norms agree (orthonormal real basis)
    xr = torch.view_as_real(Xr)
    assert relerr((xr ** 2).sum(dim=(2, 3)).cpu(), (Zs.abs() ** 2).sum(dim=2).cpu()) < 1e-13
    he, ho = (m + 1) // 2, m // 2
    Ve, Vo = V[0::2, :he].contiguous(), V[1::2, :ho].contiguous()
    # separable diagonal d(i, j) = dh[i] + dc[j] (the form the LC energies really have)
    g = torch.Generator().manual_seed(77)
    dh = torch.randn(m, dtype=torch.float64, generator=g)
    dc_half = torch.randn(nc_ + 1, dtype=torch.float64, generator=g)
    dc_r = torch.cat([dc_half, dc_half[1:]])                          # real layout: cos and sin planes
    sep = torch.cat([dh, dc_r]).cuda()
    fdc_sep = (dh[:, None] + torch.cat([dc_half.flip(0)[:nc_], dc_half])[None, :]).reshape(1, N).contiguous().cuda()
    for use_fd in (True, False, 'sep'):
        Yc = torch.zeros_like(Zs)
        fd_c = fdc_sep if use_fd == 'sep' else (fdc if use_fd else None)
        assert kern.hx_fused(sp, Ve, Vo, pot_struct(shifts), fd_c, BlockView(Zs, cnt=cnt), BlockView(Yc, cnt=cnt))
        Yr = torch.zeros_like(Xr)
        assert kern.hx_fused_real(sp, Ve, Vo, pot_struct(shifts), fdr if use_fd is True else None,
                                  BlockView(Xr, cnt=cnt), BlockView(Yr, cnt=cnt),
                                  fdiag_sep=sep if use_fd == 'sep' else None)
        Yrc = torch.zeros_like(Zs)
        kern.real_to_complex(sp, BlockView(Yr, cnt=cnt), BlockView(Yrc, cnt=cnt))
        assert relerr(Yrc.cpu(), Yc.cpu()) < 2e-13, (dims, use_fd)
        if N % 2:     # the padding double of an odd N stays zero
            assert float(torch.view_as_real(Yr)[:, :, -1, 1].abs().max()) == 0.0


def test_hx_real_many_vectors(kern):
    """More vectors than resident teams: the persistent loop reuses the shared buffers (TMA store ->
    next TMA load), per-point offsets / counts."""
    m, inner = 100, 51
    N, Nc = m * inner, (m * inner + 1) // 2
    B, nv = 40, 12
    sp, V, etab, a, gl, fdc, fdr, pot_struct = _real_case(kern, m, inner, B, seed=50)
    g = torch.Generator().manual_seed(7)
    cnt = torch.randint(0, nv + 1, (B,), dtype=torch.int32, generator=g).cuda()
    off = torch.randint(0, 5, (B,), dtype=torch.int32, generator=g).cuda()
    Zs = torch.zeros((B, nv + 4, N), dtype=torch.complex128, device='cuda')
    Xr = torch.view_as_complex(torch.randn((B, nv + 4, Nc, 2), dtype=torch.float64, generator=g).cuda())
    kern.real_to_complex(sp, BlockView(Xr), BlockView(Zs))
    he, ho = (m + 1) // 2, m // 2
    Ve, Vo = V[0::2, :he].contiguous(), V[1::2, :ho].contiguous()
    Yc = torch.zeros_like(Zs)
    kern.hx_fused(sp, Ve, Vo, pot_struct(), fdc, BlockView(Zs, cnt_fixed=nv, off=off, cnt=cnt),
                  BlockView(Yc, cnt_fixed=nv, off=off, cnt=cnt))
    Yr = torch.zeros_like(Xr)
    assert kern.hx_fused_real(sp, Ve, Vo, pot_struct(), fdr, BlockView(Xr, cnt_fixed=nv, off=off, cnt=cnt),
                              BlockView(Yr, cnt_fixed=nv, off=off, cnt=cnt))
    Yrc = torch.zeros_like(Zs)
    kern.real_to_complex(sp, BlockView(Yr), BlockView(Yrc))
    assert relerr(Yrc.cpu(), Yc.cpu()) < 2e-13
    # H is real symmetric in this representation: <x, H y> = <H x, y>
    xr, yr = torch.view_as_real(Xr).reshape(B, nv + 4, -1), torch.view_as_real(Yr).reshape(B, nv + 4, -1)
    b = int(torch.argmax(cnt))
    o0, c = int(off[b]), int(cnt[b])
    G = xr[b, o0:o0 + c] @ yr[b, o0:o0 + c].T
    assert relerr(G.cpu(), G.T.cpu()) < 1e-12


def test_lowblock_assemble_real_is_the_rotated_complex_block(kern):
    """Low-energy block in the real basis = U^H (complex block on the +-n closure of the set) U."""
    m, inner, B = 25, 13, 2
    nc = (inner - 1) // 2
    N = m * inner
    sp, V, etab, a, gl, fdc, fdr, pot_struct = _real_case(kern, m, inner, B, seed=60)
    Dt = torch.einsum('ni,ti,ki->tnk', V.to(torch.complex128), etab[0, :, :m], V.to(torch.complex128)).contiguous()
    units = [(i, n) for i in range(6) for n in range(4)]          # whole (Fock, |charge|) units
    Sr, Sc = [], []
    for i, n in units:
        Sr.append(i * inner + n)
        Sc.append(i * inner + nc + n)
        if n:
            Sr.append(i * inner + nc + n)
            Sc.append(i * inner + nc - n)
    ns = len(Sr)
    U = np.zeros((ns, ns), dtype=complex)       # columns: real basis states in terms of the complex ones
    r = 0
    for i, n in units:
        if n == 0:
            U[r, r] = 1.0
            r += 1
        else:           # complex order (+n, -n); real order (cos, sin): c = (|n>+|-n>)/sqrt2, s = i(|n>-|-n>)/sqrt2
            U[r, r], U[r + 1, r] = 2 ** -0.5, 2 ** -0.5
            U[r, r + 1], U[r + 1, r + 1] = 1j * 2 ** -0.5, -1j * 2 ** -0.5
            r += 2
    sigma = torch.tensor([-3.0, -4.0], dtype=torch.float64, device='cuda')
    Srd = torch.tensor([Sr], dtype=torch.int32, device='cuda')
    Scd = torch.tensor([Sc], dtype=torch.int32, device='cuda')
    outc = torch.empty((B, ns, ns), dtype=torch.complex64, device='cuda')
    kern.lowblock_assemble(sp, Scd, Dt, pot_struct(), fdc, sigma, 1.0, outc)
    outr = torch.empty((B, ns, ns), dtype=torch.complex64, device='cuda')
    kern.lowblock_assemble(sp, Srd, Dt, pot_struct(), fdr, sigma, 1.0, outr, real=True)
    Ut = torch.as_tensor(U, device='cuda')
    # the real-representation amplitudes are y = U^H z (y = sqrt2 Re z, sqrt2 Im z): H_r = U^H H_c U ... with the
    # kernel's convention z = U y
    ref = Ut.conj().T @ outc.to(torch.complex128) @ Ut
    assert float(ref.imag.abs().max()) < 1e-5 * float(ref.abs().max())
    assert relerr(outr.to(torch.complex128).cpu(), ref.cpu()) < 1e-6
    assert float(outr.imag.abs().max()) == 0.0
    outf = torch.empty((B, ns, ns), dtype=torch.float32, device='cuda')         # float storage (banded kernels)
    kern.lowblock_assemble(sp, Srd, Dt, pot_struct(), fdr, sigma, 1.0, outf, real=True)
    assert torch.equal(outf, outr.real)


@pytest.mark.parametrize('mm,sizes', [(40, [40, 36, 24, 12, 1, 0, 33]), (64, [64, 61, 12]), (16, [16, 7])])
def test_gen_rr_real(kern, mm, sizes):
    """Real symmetric generalised Rayleigh-Ritz (128-thread CTA per pencil) against scipy: eigenvalues,
    GB-orthonormal coefficient vectors, dependent basis vector removed, only the n_out lowest vectors
    written."""
    import scipy.linalg
    B, n_out = len(sizes), 12
    g = torch.Generator().manual_seed(51)
    Q = torch.randn((B, mm, 150), dtype=torch.float64, generator=g)
    if sizes[1] > 8:
        Q[1, 7] = 0.5 * Q[1, 3] - 0.25 * Q[1, 5]          # dependent -> must be removed
    Hs = torch.randn((B, 150, 150), dtype=torch.float64, generator=g)
    Hs = Hs + Hs.transpose(1, 2)
    GB = (Q @ Q.transpose(1, 2)).to(torch.complex128)
    GA = (Q @ Hs @ Q.transpose(1, 2)).to(torch.complex128)
    # the imaginary parts of a real-representation Gram matrix are meaningless: must be ignored
    GBd = GB + 1j * torch.randn((B, mm, mm), dtype=torch.float64, generator=g)
    GAd = GA + 1j * torch.randn((B, mm, mm), dtype=torch.float64, generator=g)
    nn = torch.tensor(sizes, dtype=torch.int32)
    w = torch.full((B, mm), -7.0, dtype=torch.float64, device='cuda')
    Zt = torch.full((B, mm, mm), 3.0 + 2.0j, dtype=torch.complex128, device='cuda')
    kern.gen_rr(GAd.cuda(), GBd.cuda(), nn.cuda(), 1e-8, w, Zt, real=True, n_out=n_out)
    w, Zt = w.cpu(), Zt.cpu()
    for b in range(B):
        nb = sizes[b]
        if nb == 0:
            assert float(w[b, 0]) == -7.0
            continue
        a, gm = GA[b, :nb, :nb].real.numpy(), GB[b, :nb, :nb].real.numpy()
        keep = [i for i in range(nb) if not (b == 1 and sizes[1] > 8 and i == 7)]
        wr = scipy.linalg.eigh(a[np.ix_(keep, keep)], gm[np.ix_(keep, keep)], eigvals_only=True)
        k = len(wr)
        assert np.abs(w[b, :k].numpy() - wr).max() < 1e-10 * np.abs(wr).max(), (b, nb)
        assert np.all(np.isinf(w[b, k:].numpy()))
        ko = min(k, n_out)
        c = Zt[b, :ko, :nb].numpy().T
        assert np.abs(c.imag).max() == 0.0
        c = c.real
        assert np.abs(c.T @ gm @ c - np.eye(ko)).max() < 1e-9
        assert np.abs(a @ c - gm @ c * w[b, :ko].numpy()[None, :]).max() < 1e-9 * np.abs(wr).max() * np.abs(c).max()
        if nb > n_out:      # rows beyond n_out are not written
            assert complex(Zt[b, n_out, 0]) == 3.0 + 2.0j


def test_lowband_real_factor_and_apply(kern):
    """Float-storage banded preconditioner of the real representation: block LDL^T factor and substitution
    (four right-hand sides per warp) against the emulation and a dense solve; real amplitudes are gathered
    from / scattered to blocks of complex elements."""
    rng = np.random.default_rng(9)
    sizes = [[3, 7, 12, 33, 40, 12, 6, 3], [5, 5, 50, 36, 20], [64, 52]]
    B, ns, Nc, p = 5, 116, 400, 7
    nslot, group = 3, 2
    off = np.zeros((nslot, 9), dtype=np.int32)
    nblk = np.zeros(nslot, dtype=np.int32)
    A = np.zeros((nslot, ns, ns))
    for sl, sz in enumerate(sizes):
        assert sum(sz) == ns
        o = np.concatenate(([0], np.cumsum(sz)))
        off[sl, :len(o)] = o
        nblk[sl] = len(sz)
        for q in range(len(sz)):
            a, b_ = o[q], o[q + 1]
            d = rng.standard_normal((sz[q], sz[q]))
            A[sl, a:b_, a:b_] = d @ d.T / sz[q] + 4 * np.eye(sz[q])
            if q + 1 < len(sz):
                c = o[q + 2]
                e = 0.3 * rng.standard_normal((sz[q], sz[q + 1]))
                A[sl, a:b_, b_:c] = e
                A[sl, b_:c, a:b_] = e.T
    A32 = torch.from_numpy(A.astype(np.float32))
    offt, nbt = torch.from_numpy(off), torch.from_numpy(nblk)
    E = EmulKernels()
    fref = A32.clone()
    E.lowband_factor(fref, offt, nbt, 64)
    fac = A32.clone().cuda()
    kern.lowband_factor(fac, offt.cuda(), nbt.cuda(), 64)
    assert relerr(fac.cpu().double(), fref.double()) < 2e-5
    S = torch.stack([torch.randperm(2 * Nc, generator=torch.Generator().manual_seed(s))[:ns] for s in range(nslot)]).to(torch.int32)
    X, HX = rnd((B, p, Nc), 61), rnd((B, p, Nc), 62)
    Vb = rnd((B, 20, Nc), 63)
    theta = torch.randn((B, 48), dtype=torch.float64, generator=torch.Generator().manual_seed(64))
    msz = torch.tensor([6, 10, 13, 3, 8], dtype=torch.int32)
    nact = torch.tensor([7, 2, 0, 5, 1], dtype=torch.int32)
    act = torch.tensor([[0, 1, 2, 3, 4, 5, 6], [4, 1, 0, 0, 0, 0, 0], [0] * 7, [5, 3, 2, 0, 6, 0, 0], [2, 0, 0, 0, 0, 0, 0]],
                       dtype=torch.int32)
    refV = Vb.clone()
    E.lowband_apply(BlockView(X), BlockView(HX), BlockView(refV, cnt_fixed=p, off=msz, cnt=nact), theta, act, S,
                    fref, offt, nbt, 64, 2.0, group=group, real=True)
    for perm in (None, torch.argsort(S, dim=1).to(torch.int32).cuda()):
        Vd = Vb.cuda()
        kern.lowband_apply(BlockView(X.cuda()), BlockView(HX.cuda()),
                           BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()), theta.cuda(), act.cuda(),
                           S.cuda(), fac, offt.cuda(), nbt.cuda(), 64, 2.0, group=group, perm=perm, real=True)
        assert relerr(torch.view_as_real(Vd.cpu()), torch.view_as_real(refV)) < 2e-5
    b, jj = 0, 1
    s = S[0].long().numpy()
    xr, hr = torch.view_as_real(X)[b, jj].reshape(-1).numpy(), torch.view_as_real(HX)[b, jj].reshape(-1).numpy()
    sol = np.linalg.solve(A[0], (hr[s] - float(theta[b, jj]) * xr[s]) / 2.0)
    got = torch.view_as_real(Vd.cpu())[b, int(msz[b]) + jj].reshape(-1).numpy()[s]
    assert np.abs(got - sol).max() / np.abs(sol).max() < 1e-4
    with pytest.raises(ValueError):
        kern.lowband_apply(BlockView(X.cuda()), BlockView(HX.cuda()),
                           BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()), theta.cuda(), act.cuda(),
                           S.cuda(), fac, offt.cuda(), nbt.cuda(), 64, 2.0, group=group, real=False)


@pytest.mark.parametrize('B,mm,p,N', [(5, 48, 12, 3001), (3, 40, 12, 5050), (300, 24, 8, 777), (4, 16, 16, 130), (2, 64, 5, 64)])
def test_gram_pair_real(kern, B, mm, p, N):
    """Paired Gram pass of the real representation (eight right-hand vectors per DMMA column tile): real
    inner products over the 2 N doubles of every vector, zero imaginary parts."""
    g = torch.Generator().manual_seed(71)
    V, W = rnd((B, mm, N), 72), rnd((B, mm, N), 73)
    msz = torch.randint(0, mm - p + 1, (B,), dtype=torch.int32, generator=g)
    nact = torch.randint(0, p + 1, (B,), dtype=torch.int32, generator=g)
    msz[0], nact[0] = mm - p, p
    mtot = msz + nact
    Vd, Wd = V.cuda(), W.cuda()
    o1 = torch.full((B, mm, 16), 3.0 + 1j, dtype=torch.complex128, device='cuda')
    o2 = torch.full((B, mm, 16), 3.0 + 1j, dtype=torch.complex128, device='cuda')
    kern.gram_pair(BlockView(Vd, cnt_fixed=mm, cnt=mtot.cuda()), BlockView(Vd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()),
                   BlockView(Wd, cnt_fixed=p, off=msz.cuda(), cnt=nact.cuda()), o1, o2, real=True)
    Vr, Wr = torch.view_as_real(V).reshape(B, mm, 2 * N), torch.view_as_real(W).reshape(B, mm, 2 * N)
    for b in range(B):
        a = Vr[b, :int(mtot[b])]
        t, ht = Vr[b, int(msz[b]):int(mtot[b])], Wr[b, int(msz[b]):int(mtot[b])]
        na = int(nact[b])
        if na == 0 or int(mtot[b]) == 0:
            continue
        r1, r2 = a @ t.T, a @ ht.T
        got1, got2 = o1[b, :int(mtot[b]), :na].cpu(), o2[b, :int(mtot[b]), :na].cpu()
        assert float(got1.imag.abs().max()) == 0.0 and float(got2.imag.abs().max()) == 0.0
        assert relerr(got1.real, r1) < 1e-13 and relerr(got2.real, r2) < 1e-13, b


# ---- sqc_hx_generic: one-kernel H.X for leading harmonic modes, against a plain torch fp64 restatement -------
def _hx_torch_reference(dims, harm, Vs, lam_list, s, shifts, a, g, fd, X):
    """Y = fd .* X + (x V_i) P ((x V_i^T) X) with dense torch tensor contractions (CPU, fp64)."""
    B, nv, N = X.shape
    M = len(dims)
    Z = X.reshape((B, nv) + tuple(dims))
    for i in range(M):
        if harm[i] and dims[i] > 1 and i in Vs:
            Z = torch.movedim(torch.tensordot(Z, Vs[i].to(torch.complex128), dims=([2 + i], [0])), -1, 2 + i)
    out = torch.zeros_like(Z)
    lin = g[:, 0].reshape((B, 1) + (1,) * M).to(torch.complex128)
    for i in range(M):
        if harm[i]:
            shp = [1] * (2 + M)
            shp[2 + i] = dims[i]
            lin = lin + g[:, 1 + i].reshape((B, 1) + (1,) * M) * lam_list[i].reshape(shp)
    out = out + lin * Z
    T = a.shape[1]
    for t in range(T):
        ph = a[:, t].reshape((B, 1) + (1,) * M)
        for i in range(M):
            if harm[i]:
                shp = [1] * (2 + M)
                shp[2 + i] = dims[i]
                sb = s[:, t, i].reshape((-1, 1) + (1,) * M)
                ph = ph * torch.exp(1j * sb * lam_list[i].reshape(shp))
        # z(i - sh): roll forward by sh with zero fill; z(i + sh): roll back
        lo, hi = Z, Z
        for i in range(M):
            w = int(shifts[t][i]) if not harm[i] else 0
            if w:
                lo = _shift(lo, 2 + i, w)
                hi = _shift(hi, 2 + i, -w)
        out = out + ph * lo + ph.conj() * hi
    for i in range(M):
        if harm[i] and dims[i] > 1 and i in Vs:
            out = torch.movedim(torch.tensordot(out, Vs[i].T.to(torch.complex128), dims=([2 + i], [0])), -1, 2 + i)
    Y = out.reshape(B, nv, N)
    if fd is not None:
        Y = Y + fd.reshape(-1, 1, N) * X
    return Y


def _shift(Z, axis, w):
    """out[..., i, ...] = Z[..., i - w, ...] (zero where i - w is out of range)."""
    out = torch.zeros_like(Z)
    n = Z.shape[axis]
    if abs(w) >= n:
        return out
    src = [slice(None)] * Z.dim()
    dst = [slice(None)] * Z.dim()
    if w > 0:
        src[axis], dst[axis] = slice(0, n - w), slice(w, n)
    else:
        src[axis], dst[axis] = slice(-w, n), slice(0, n + w)
    out[tuple(dst)] = Z[tuple(src)]
    return out


@pytest.mark.parametrize('dims,harm,shifts,per_point', [
    ([10, 9, 9], [1, 0, 0], [[0, 1, 0], [0, 0, 1]], False),                 # small transmon-resonator-transmon
    ([10, 10, 9, 9], [1, 1, 0, 0], [[0, 0, 1, 0], [0, 0, 1, 0], [0, 0, 0, 1]], True),   # the 4-mode batch layout
    ([10, 50, 50], [1, 0, 0], [[0, 1, 0], [0, 0, -1], [0, 1, 1]], False),  # column tiles with a halo of 51
    ([3, 1, 7], [1, 1, 0], [[0, 0, 1]], False),                             # a harmonic mode truncated to one level
    ([24, 5], [1, 0], [[0, 1]], False),                                     # long fibres (split re / im path)
    ([32, 3], [1, 0], [[0, -1]], True),
    ([6, 4, 5, 3, 3], [1, 1, 1, 0, 0], [[0, 0, 0, 1, -1], [0, 0, 0, 0, 1]], False)])
def test_hx_generic_matches_torch_reference(kern, dims, harm, shifts, per_point):
    from sqcircuit_b200._lib import Potential
    M, N = len(dims), int(np.prod(dims))
    B, nv, T = 3, 5, len(shifts)
    sp = make_space(dims, harm)
    toff = np.concatenate(([0], np.cumsum(dims)))[:-1]
    tlen = int(np.sum(dims))
    lam = torch.zeros(tlen, dtype=torch.float64)
    Vs, lam_list = {}, {}
    for i in range(M):
        if harm[i] and dims[i] > 1:
            Vd, ld = kern.xbasis(dims[i])
            Vs[i], lam_list[i] = Vd.cpu(), ld.cpu()
            lam[toff[i]:toff[i] + dims[i]] = ld.cpu()
        else:
            lam_list[i] = torch.zeros(dims[i], dtype=torch.float64)
    gen = torch.Generator().manual_seed(11)
    nb = B if per_point else 1
    s = 0.4 * torch.randn((nb, T, M), dtype=torch.float64, generator=gen) * torch.tensor(harm, dtype=torch.float64)
    etab = torch.empty((nb, T, tlen), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lam.cuda(), s.cuda().contiguous(), etab)
    a = rnd((B, T), 12)
    g = torch.randn((B, 1 + M), dtype=torch.float64, generator=gen)
    fd = torch.randn((B, N), dtype=torch.float64, generator=gen)
    X = rnd((B, nv, N), 13)
    cnt = torch.tensor([nv, 2, 0], dtype=torch.int32)
    Vcat = torch.cat([Vs[i].reshape(-1) for i in range(M) if i in Vs]).cuda().contiguous()
    pot = Potential()
    pot.n_terms = T
    for t in range(T):
        for i in range(M):
            pot.shift[t][i] = int(shifts[t][i])
    lamd, ad, gd = lam.cuda(), a.cuda().contiguous(), g.cuda().contiguous()
    pot.lam, pot.etab, pot.etab_stride_b = lamd.data_ptr(), etab.data_ptr(), (etab.stride(0) if per_point else 0)
    pot.g, pot.a, pot.diag, pot.diag_stride_b = gd.data_ptr(), ad.data_ptr(), None, 0
    for use_fd in (True, False):
        ref = _hx_torch_reference(dims, harm, Vs, lam_list, s, shifts, a, g, fd if use_fd else None, X)
        out = torch.full_like(X, 7.0).cuda()
        ok = kern.hx_generic(sp, Vcat, pot, fd.cuda() if use_fd else None, BlockView(X.cuda(), cnt=cnt.cuda()),
                             BlockView(out, cnt=cnt.cuda()))
        assert ok
        out = out.cpu()
        for b in range(B):
            n = int(cnt[b])
            if n:
                assert relerr(out[b, :n], ref[b, :n]) < 1e-13, (b, use_fd)
            assert bool((out[b, n:] == 7.0).all())            # vectors beyond the count are not touched
    # and the unfused kernel sequence the operator falls back to gives the same
    s1, s2 = torch.zeros_like(X).cuda(), torch.zeros_like(X).cuda()
    src = BlockView(X.cuda())
    for i in range(M):
        if i in Vs:
            kern.mode_transform(sp, i, Vs[i].cuda(), 1, src, BlockView(s1), impl=1)
            src = BlockView(s1)
    kern.potential_apply(sp, pot, BlockView(s1), BlockView(s2))
    for i in range(M):
        if i in Vs:
            kern.mode_transform(sp, i, Vs[i].cuda(), 0, BlockView(s2), BlockView(s2), impl=1)
    ref2 = _hx_torch_reference(dims, harm, Vs, lam_list, s, shifts, a, g, None, X)
    assert relerr(s2.cpu(), ref2) < 1e-13


@pytest.mark.parametrize('dims,harm,shifts,nv', [
    ([100, 101], [1, 0], [[0, 1], [0, -1]], 2),
    ([10, 9, 9], [1, 0, 0], [[0, 1, 0], [0, 0, 1], [0, 1, -1]], 2),
    ([7, 5, 3, 3], [1, 1, 0, 0], [[0, 0, 1, 0]], 4),
    ([33], [1], [[0]], 1),
    ([41], [0], [[1]], 3)])
def test_potential_elements_matches_torch_reference(kern, dims, harm, shifts, nv):
    """out[b][i][j] = <z_i| P |z_j> (sqc_potential_elements) against P applied with dense torch index shifts
    (the x-basis part of _hx_torch_reference: identity transforms) and a torch inner product."""
    from sqcircuit_b200._lib import Potential
    M, N = len(dims), int(np.prod(dims))
    B, T = 5, len(shifts)
    sp = make_space(dims, harm)
    toff = np.concatenate(([0], np.cumsum(dims)))[:-1]
    tlen = int(np.sum(dims))
    gen = torch.Generator().manual_seed(3)
    lam = torch.randn(tlen, dtype=torch.float64, generator=gen)
    lam_list = {i: lam[toff[i]:toff[i] + dims[i]] * (1.0 if harm[i] else 0.0) for i in range(M)}
    s = 0.4 * torch.randn((B, T, M), dtype=torch.float64, generator=gen) * torch.tensor(harm, dtype=torch.float64)
    etab = torch.empty((B, T, tlen), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lam.cuda(), s.cuda().contiguous(), etab)
    a = rnd((B, T), 4)
    g = torch.randn((B, 1 + M), dtype=torch.float64, generator=gen)
    Z = rnd((B, nv + 1, N), 5)              # one more vector in the block than is used
    # identity "transforms": harm flags off for the reference's rotation, on for the phases / linear terms
    ref_pz = _hx_torch_reference(dims, harm, {}, lam_list, s, shifts, a, g, None, Z[:, :nv].contiguous())
    ref = torch.einsum('bin,bjn->bij', Z[:, :nv].conj(), ref_pz)
    pot = Potential()
    pot.n_terms = T
    for t in range(T):
        for i in range(M):
            pot.shift[t][i] = int(shifts[t][i])
    lamd, ad, gd = lam.cuda(), a.cuda().contiguous(), g.cuda().contiguous()
    pot.lam, pot.etab, pot.etab_stride_b = lamd.data_ptr(), etab.data_ptr(), etab.stride(0)
    pot.g, pot.a, pot.diag, pot.diag_stride_b = gd.data_ptr(), ad.data_ptr(), None, 0
    out = torch.zeros((B, nv, nv), dtype=torch.complex128, device='cuda')
    kern.potential_elements(sp, pot, BlockView(Z.cuda()), out)
    assert relerr(out.cpu(), ref) < 1e-13
    # and the same numbers through sqc_potential_apply + sqc_gram
    S = torch.zeros_like(Z).cuda()
    kern.potential_apply(sp, pot, BlockView(Z.cuda()), BlockView(S))
    assert relerr(out.cpu(), torch.einsum('bin,bjn->bij', Z[:, :nv].conj(), S.cpu()[:, :nv])) < 1e-13


@pytest.mark.parametrize('dims,harm,shifts', [
    ([6, 5, 5], [1, 0, 0], [[0, 1, 0], [0, 0, 1]]),
    ([5, 4, 3, 3], [1, 1, 0, 0], [[0, 0, 1, 0], [0, 0, 1, -1], [0, 0, 0, 1]]),
    ([7, 1, 5], [1, 1, 0], [[0, 0, 1]]),
    ([12, 9], [1, 0], [[0, 1], [0, -1]]),
    ([9, 6], [1, 1], [[0, 0], [0, 0]])])
def test_lowblock_assemble_generic_matches_dense_operator(kern, dims, harm, shifts):
    """(H_SS - sigma) / unit from the per-mode displacement tables (sqc_lowblock_assemble_generic) against the dense
    matrix of the same operator -- H applied to the identity by the transform / potential kernels -- restricted to
    the index set; per-point phase tables, index sets and coefficients."""
    from sqcircuit_b200._lib import Potential
    M, N = len(dims), int(np.prod(dims))
    B, T, ns = 3, len(shifts), min(40, N // 2)
    sp = make_space(dims, harm)
    toff = np.concatenate(([0], np.cumsum(dims)))[:-1]
    tlen = int(np.sum(dims))
    lam = torch.zeros(tlen, dtype=torch.float64, device='cuda')
    Vs = {}
    for i in range(M):
        if harm[i] and dims[i] > 1:
            Vs[i], li = kern.xbasis(dims[i])
            lam[toff[i]:toff[i] + dims[i]] = li
    gen = torch.Generator().manual_seed(21)
    s = (0.5 * torch.randn((B, T, M), dtype=torch.float64, generator=gen) * torch.tensor(harm, dtype=torch.float64)).cuda()
    etab = torch.empty((B, T, tlen), dtype=torch.complex128, device='cuda')
    kern.phase_tables(sp, T, lam, s.contiguous(), etab)
    a = rnd((B, T), 22).cuda()
    g = torch.randn((B, 1 + M), dtype=torch.float64, generator=gen)
    for i in range(M):
        if not harm[i] or dims[i] == 1:
            g[:, 1 + i] = 0.0                    # no linear term on charge modes / one-level modes (host contract)
    g = g.cuda()
    d = torch.randn((B, N), dtype=torch.float64, generator=gen).cuda()
    pot = Potential()
    pot.n_terms = T
    for t in range(T):
        for i in range(M):
            pot.shift[t][i] = int(shifts[t][i])
    pot.lam, pot.etab, pot.etab_stride_b = lam.data_ptr(), etab.data_ptr(), etab.stride(0)
    pot.g, pot.a, pot.diag, pot.diag_stride_b = g.data_ptr(), a.data_ptr(), None, 0
    # dense H of every point: operator applied to the identity (column j of H = H e_j, stored as row j)
    eye = torch.eye(N, dtype=torch.complex128, device='cuda').unsqueeze(0).expand(B, N, N).contiguous()
    s1, s2 = torch.zeros_like(eye), torch.zeros_like(eye)
    src = BlockView(eye)
    for i in Vs:
        kern.mode_transform(sp, i, Vs[i], 1, src, BlockView(s1), impl=1)
        src = BlockView(s1)
    kern.potential_apply(sp, pot, src, BlockView(s2))
    for i in Vs:
        kern.mode_transform(sp, i, Vs[i], 0, BlockView(s2), BlockView(s2), impl=1)
    H = s2.transpose(1, 2) + torch.diag_embed(d.to(torch.complex128))          # [B][N][N]
    # tables, index sets, shifts
    parts = []
    for i in Vs:
        m, o = dims[i], int(toff[i])
        V = Vs[i].to(torch.complex128)
        parts.append(torch.einsum('ni,bti,ki->btnk', V, etab[:, :, o:o + m], V).reshape(B, T, m * m))
    Dt = torch.cat(parts, dim=2).contiguous()
    S = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(30 + b))[:ns] for b in range(B)]).to(torch.int32).cuda()
    sigma = torch.tensor([-3.0, 0.5, 2.0], dtype=torch.float64, device='cuda')
    unit = 1.7
    out = torch.zeros((B, ns, ns), dtype=torch.complex64, device='cuda')
    kern.lowblock_assemble_generic(sp, S, Dt, pot, d, sigma, unit, out)
    for b in range(B):
        idx = S[b].long()
        ref = (H[b][idx][:, idx] - sigma[b] * torch.eye(ns, device='cuda')) / unit
        assert relerr(out[b].to(torch.complex128).cpu(), ref.cpu()) < 1e-6           # complex64 storage
    # the whole matrix in complex128 (sqc_hamiltonian_dense: the dense eigensolver path)
    Hd = torch.empty((B, N, N), dtype=torch.complex128, device='cuda')
    kern.hamiltonian_dense(sp, Dt, pot, d, Hd)
    assert relerr(Hd.cpu(), H.cpu()) < 1e-13
    if all(harm):
        # no charge modes: H is real symmetric in the Fock basis, real output for sqc_syev
        assert float(H.imag.abs().max()) < 1e-13 * float(H.abs().max())
        Hr = torch.empty((B, N, N), dtype=torch.float64, device='cuda')
        kern.hamiltonian_dense(sp, Dt, pot, d, Hr)
        assert relerr(Hr.cpu().to(torch.complex128), H.cpu()) < 1e-13
    # the emulated contract gives the same block
    pot._keep = (a.cpu(), g.cpu(), etab.cpu(), None)
    oute = torch.zeros((B, ns, ns), dtype=torch.complex64)
    EmulKernels().lowblock_assemble_generic(sp, S.cpu(), Dt.cpu(), pot, d.cpu(), sigma.cpu(), unit, oute)
    assert relerr(out.cpu().to(torch.complex128), oute.to(torch.complex128)) < 1e-6


def test_pinned_staging_upload_download_and_free_memory_tracking(kern):
    """Kernels.upload / download (pinned staging instead of pageable driver copies) round-trip every dtype the
    host code sends, survive the wrap-around of the staging arena, fall back to a plain copy for large arrays;
    free_bytes follows the allocator between its rare cudaMemGetInfo calls."""
    rng = np.random.default_rng(5)
    for arr in (rng.standard_normal((3, 7)), rng.integers(0, 1000, 33), rng.standard_normal(5) + 1j * rng.standard_normal(5),
                np.arange(12, dtype=np.int32).reshape(3, 4)[:, ::2], np.broadcast_to(np.arange(4.0), (3, 4)), np.zeros((0,))):
        t = kern.upload(arr)
        assert t.device.type == 'cuda' and tuple(t.shape) == arr.shape
        assert np.array_equal(kern.download(t), np.asarray(arr))
    assert kern.upload(np.arange(5), torch.float64).dtype == torch.float64
    big = rng.standard_normal(kern._PIN_BYTES // 8)                      # > a quarter of the arena: plain copy
    assert np.array_equal(kern.download(kern.upload(big)), big)
    chunk = rng.standard_normal(kern._PIN_BYTES // 40)                   # forces several wrap-arounds
    outs = [kern.upload(chunk + i) for i in range(30)]
    for i, t in enumerate(outs):
        assert np.array_equal(t.cpu().numpy(), chunk + i)
    f0 = kern.free_bytes()
    hold = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
    f1 = kern.free_bytes()
    assert f1 <= f0 and f0 - f1 <= (1 << 30) + (64 << 20)                # the new GiB is held; the rest is slack
    del hold
