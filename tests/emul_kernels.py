"""TEST-ONLY numpy emulation of the kernel contracts in include/sqcircuit_b200.h.

The product has exactly one backend (the CUDA library).  This module restates what each C
entry point is specified to compute so that the *host logic* (operator tables, block-Davidson
orchestration, decoherence formulas) can be checked on a machine without a GPU
(`pytest -m "not gpu"`).  It is injected through ``Circuit(..., kernels=EmulKernels())``; nothing
under ``sqcircuit_b200/`` imports it.
"""
import numpy as np
import torch


def _cnt(blk, b):
    return int(blk.cnt[b]) if blk.cnt is not None else blk.cnt_fixed


def _off(blk, b):
    return blk.first + (int(blk.off[b]) if blk.off is not None else 0)


def _bi(blk, b):
    """index of sweep point b in the tensor (0 for a shared pool, stride_b == 0)"""
    return 0 if getattr(blk, 'stride_b', None) == 0 else b


def _get(blk, b):
    o, c = _off(blk, b), _cnt(blk, b)
    return blk.tensor[_bi(blk, b), o:o + c, :].numpy()


def _rows(blk, b):
    """all cnt_fixed rows of the block of sweep point b (Ritz index addressing)"""
    o = _off(blk, b)
    return blk.tensor[_bi(blk, b), o:o + blk.cnt_fixed, :].numpy()


def _set(blk, b, arr):
    o = _off(blk, b)
    blk.tensor[_bi(blk, b), o:o + arr.shape[0], :] = torch.from_numpy(np.ascontiguousarray(arr))


class EmulKernels:
    name = 'emul'

    def __init__(self):
        self.dev = torch.device('cpu')
        self.launches = 0

    def workspace(self, key, shape, dtype):
        """Persistent scratch per key, like Kernels.workspace (callers rely on getting the same storage back)."""
        shape = tuple(int(d) for d in shape)
        ws = self.__dict__.setdefault('_ws', {})
        n = int(np.prod(shape))
        buf = ws.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            ws[key] = buf = torch.empty(n, dtype=dtype)
        return buf[:n].view(*shape)

    def release_workspace(self):
        self.__dict__['_ws'] = {}

    free_bytes_value = 1 << 60     # tests shrink it to exercise the memory guard of Circuit.sweep_diag

    def free_bytes(self):
        return self.free_bytes_value

    def launch_count(self):
        return self.launches

    # ---- (1) --------------------------------------------------------------------------
    def xbasis(self, m):
        x = np.diag(np.sqrt(np.arange(1, m)), 1)
        x = x + x.T
        lam, V = np.linalg.eigh(x)
        for j in range(m):
            nz = np.where(np.abs(V[:, j]) > 1e-8)[0][0]
            if V[nz, j] < 0:
                V[:, j] = -V[:, j]
        return torch.from_numpy(V.copy()), torch.from_numpy(lam.copy())

    def phase_tables(self, sp, n_terms, lam, s, etab):
        dims = [sp.dims[i] for i in range(sp.n_modes)]
        off = np.concatenate(([0], np.cumsum(dims)))
        lamn = lam.numpy()
        for b in range(etab.shape[0]):
            sb = s[min(b, s.shape[0] - 1)].numpy()
            for t in range(n_terms):
                for m in range(sp.n_modes):
                    seg = slice(off[m], off[m + 1])
                    if sp.harmonic[m]:
                        etab[b, t, seg] = torch.from_numpy(np.exp(1j * sb[t, m] * lamn[seg]))
                    else:
                        etab[b, t, seg] = 1.0

    def lc_diagonal(self, sp, omega, q, ng, out):
        M = sp.n_modes
        dims = [sp.dims[i] for i in range(M)]
        grids = np.meshgrid(*[np.arange(d) for d in dims], indexing='ij')
        for b in range(out.shape[0]):
            val = []
            for m in range(M):
                if sp.harmonic[m]:
                    val.append(grids[m].astype(float))
                else:
                    val.append(grids[m] - (dims[m] - 1) // 2 - float(ng[b, m]))
            acc = np.zeros(dims)
            for m in range(M):
                if sp.harmonic[m]:
                    acc += float(omega[b, m]) * val[m]
                else:
                    for m2 in range(m, M):
                        if not sp.harmonic[m2]:
                            acc += float(q[b, m, m2]) * val[m] * val[m2]
            out[b] = torch.from_numpy(acc.reshape(-1))

    # ---- (2) --------------------------------------------------------------------------
    def diag_mul(self, X, Y, diag, accumulate):
        for b in range(X.B):
            d = diag[min(b, diag.shape[0] - 1)].numpy()
            r = _get(X, b) * d[None, :]
            if accumulate:
                r = r + _get(Y, b)[:r.shape[0]]
            _set(Y, b, r)

    def mode_transform(self, sp, mode, mat, transpose, X, Z, ediag=None, EX=None, impl=0):
        dims = [sp.dims[i] for i in range(sp.n_modes)]
        for b in range(X.B):
            x = _get(X, b)
            if x.shape[0] == 0:
                continue
            m = mat.numpy() if mat.dim() == 2 else mat[min(b, mat.shape[0] - 1)].numpy()
            mm = m.T if transpose else m
            t = x.reshape([x.shape[0]] + dims)
            t = np.moveaxis(np.tensordot(mm, t, axes=([1], [1 + mode])), 0, 1 + mode)
            r = t.reshape(x.shape[0], -1)
            if ediag is not None:
                r = r + ediag[min(b, ediag.shape[0] - 1)].numpy()[None, :] * _get(EX, b)
            _set(Z, b, r)

    def potential_elements(self, sp, pot, Z, out):
        """out[b][i][j] = <z_i| P |z_j>: the contract of sqc_potential_elements, through the emulated apply."""
        from sqcircuit_b200.kernels import BlockView
        nv = out.shape[1]
        S = torch.zeros_like(Z.tensor)
        self.potential_apply(sp, pot, Z, BlockView(S))
        for b in range(Z.B):
            z = _get(Z, b)[:nv]
            s = _get(BlockView(S), b)[:nv]
            out[b] = torch.as_tensor(z.conj() @ s.T)

    def potential_apply(self, sp, pot, X, Z):
        a, g, etab, diag = pot._keep
        M = sp.n_modes
        dims = [sp.dims[i] for i in range(M)]
        off = np.concatenate(([0], np.cumsum(dims)))
        lam = torch.frombuffer(bytearray(0), dtype=torch.float64) if False else None
        lamn = self._lam_from_ptr(pot, off[-1])
        for b in range(X.B):
            x = _get(X, b)
            nv = x.shape[0]
            if nv == 0:
                continue
            t = x.reshape([nv] + dims)
            gb = g[b].numpy()
            lin = np.full(dims, gb[0])
            for m in range(M):
                if sp.harmonic[m]:
                    shp = [1] * M
                    shp[m] = dims[m]
                    lin = lin + gb[1 + m] * lamn[off[m]:off[m + 1]].reshape(shp)
            if diag is not None:
                lin = lin + diag[min(b, diag.shape[0] - 1)].numpy().reshape(dims)
            out = lin[None] * t
            eb = etab[min(b, etab.shape[0] - 1)].numpy()
            for tt in range(pot.n_terms):
                ph = np.full(dims, complex(a[b, tt]))
                sh = [0] * M
                for m in range(M):
                    if sp.harmonic[m]:
                        shp = [1] * M
                        shp[m] = dims[m]
                        ph = ph * eb[tt, off[m]:off[m + 1]].reshape(shp)
                    else:
                        sh[m] = int(pot.shift[tt][m])
                out = out + ph[None] * self._shifted(t, sh, +1) + np.conj(ph)[None] * self._shifted(t, sh, -1)
            _set(Z, b, out.reshape(nv, -1))

    @staticmethod
    def _lam_from_ptr(pot, n):
        import ctypes
        arr = (ctypes.c_double * int(n)).from_address(pot.lam)
        return np.frombuffer(arr, dtype=np.float64).copy()

    @staticmethod
    def _shifted(t, sh, sign):
        """out(i) = t(i - sign*sh) with zero fill outside the truncated space."""
        out = t
        for m, s in enumerate(sh):
            s = sign * s
            if s == 0:
                continue
            ax = 1 + m
            res = np.zeros_like(out)
            n = out.shape[ax]
            src = [slice(None)] * out.ndim
            dst = [slice(None)] * out.ndim
            if s > 0:
                dst[ax], src[ax] = slice(s, n), slice(0, n - s)
            else:
                dst[ax], src[ax] = slice(0, n + s), slice(-s, n)
            res[tuple(dst)] = out[tuple(src)]
            out = res
        return out

    def ladder_apply(self, sp, mode, c_dn, c_up, X, Y, accumulate):
        dims = [sp.dims[i] for i in range(sp.n_modes)]
        m = dims[mode]
        a = np.diag(np.sqrt(np.arange(1, m)), 1)
        for b in range(X.B):
            x = _get(X, b)
            op = complex(c_dn[b]) * a + complex(c_up[b]) * a.T
            t = x.reshape([x.shape[0]] + dims)
            r = np.moveaxis(np.tensordot(op, t, axes=([1], [1 + mode])), 0, 1 + mode).reshape(x.shape[0], -1)
            if accumulate:
                r = r + _get(Y, b)[:r.shape[0]]
            _set(Y, b, r)

    # ---- (3) --------------------------------------------------------------------------
    def gram(self, A, Bm, Cout, impl=0):
        B, ldr, ldc = Cout.shape
        rows, cols = min(A.cnt_fixed, ldr), min(Bm.cnt_fixed, ldc)
        for b in range(B):
            a, c = _get(A, b), _get(Bm, b)
            Cout[b, :rows, :cols] = 0
            if a.shape[0] and c.shape[0]:
                Cout[b, :a.shape[0], :c.shape[0]] = torch.from_numpy(a.conj() @ c.T)

    def gram_pair(self, A, Bm, Bm2, C1, C2, real=False):
        self.gram(A, Bm, C1)
        self.gram(A, Bm2, C2)
        if real:          # real representation: only the real inner products exist
            C1.imag.zero_()
            C2.imag.zero_()

    def gen_rr(self, GA, GB, n, drop, w, Zt, real=False, n_out=None):
        B, ld, _ = GA.shape
        for b in range(B):
            nb = int(n[b])
            if nb <= 0:
                continue
            a = GA[b, :nb, :nb].numpy().copy()
            g = GB[b, :nb, :nb].numpy().copy()
            if real:
                a, g = a.real.astype(complex), g.real.astype(complex)
            a = (a + a.conj().T) / 2
            g = (g + g.conj().T) / 2
            gd = np.real(np.diag(g)).copy()
            # Cholesky with dead-column handling (same rule as the kernel)
            L = np.zeros((nb, nb), complex)
            dead = np.zeros(nb, bool)
            for j in range(nb):
                d = (g[j, j] - L[j, :j] @ L[j, :j].conj()).real
                if not (d > drop * gd[j]) or not (gd[j] > 0):
                    dead[j] = True
                    L[j, :j] = 0
                    L[j, j] = 1
                    continue
                L[j, j] = np.sqrt(d)
                for i in range(j + 1, nb):
                    L[i, j] = (g[i, j] - L[i, :j] @ L[j, :j].conj()) / L[j, j]
            alive = np.where(~dead)[0]
            La = L[np.ix_(alive, alive)]
            Li = np.linalg.inv(La)
            M = Li @ a[np.ix_(alive, alive)] @ Li.conj().T
            ww, y = np.linalg.eigh((M + M.conj().T) / 2)
            c = np.zeros((nb, len(alive)), complex)
            c[alive, :] = Li.conj().T @ y
            w[b] = float('inf')
            w[b, :len(alive)] = torch.from_numpy(ww)
            rows = min(nb, n_out) if (real and n_out) else nb
            Zt[b, :rows, :nb] = 0
            keep = min(len(alive), rows)
            Zt[b, :keep, :nb] = torch.from_numpy(np.ascontiguousarray(c.T[:keep].real if real else c.T[:keep]))

    def davidson_insert2(self, st, SA, SB, B):
        for b in range(B):
            na, msz = int(st.nact[b]), int(st.msz[b])
            if na <= 0:
                continue
            for S, G in ((SA, st.G), (SB, st.GB)):
                s = S[b].numpy()
                blk = s[:msz + na, :na].copy()
                if getattr(st, 'real_basis', 0):
                    blk = blk.real.astype(complex)
                d = blk[msz:, :]
                blk[msz:, :] = (d + d.conj().T) / 2
                Gn = G[b].numpy()
                Gn[:msz + na, msz:msz + na] = blk
                Gn[msz:msz + na, :msz + na] = blk.conj().T
            st.msz[b] = msz + na

    def hx_fused(self, sp, Ve, Vo, pot, fdiag, X, Y):
        return False      # not emulated: callers fall back to transform + potential + transform

    def hx_generic(self, sp, Vcat, pot, fdiag, X, Y):
        return False      # not emulated either: same fall-back

    # ---- real (time-reversal adapted) representation ------------------------------------------
    @staticmethod
    def _r2c(y, m, inner):
        """[nv][>= m*inner] real amplitudes (planar layout) -> [nv][m*inner] complex"""
        nc = (inner - 1) // 2
        y = y[:, :m * inner].reshape(-1, m, inner)
        z = np.zeros((y.shape[0], m, inner), complex)
        z[:, :, nc] = y[:, :, 0]
        p = (y[:, :, 1:nc + 1] + 1j * y[:, :, nc + 1:]) / np.sqrt(2)
        z[:, :, nc + 1:] = p
        z[:, :, :nc] = p.conj()[:, :, ::-1]
        return z.reshape(y.shape[0], -1)

    @staticmethod
    def _c2r(z, m, inner):
        nc = (inner - 1) // 2
        z = z.reshape(-1, m, inner)
        N = m * inner
        y = np.zeros((z.shape[0], m, inner))
        y[:, :, 0] = z[:, :, nc].real
        pos, neg = z[:, :, nc + 1:], z[:, :, :nc][:, :, ::-1]
        y[:, :, 1:nc + 1] = (pos.real + neg.real) / np.sqrt(2)
        y[:, :, nc + 1:] = (pos.imag - neg.imag) / np.sqrt(2)
        out = np.zeros((z.shape[0], 2 * ((N + 1) // 2)))
        out[:, :N] = y.reshape(z.shape[0], -1)
        return out

    def real_to_complex(self, sp, Xr, Zc):
        m, inner = sp.dims[0], sp.dims[1]
        for b in range(Xr.B):
            y = _get(Xr, b)
            if y.shape[0]:
                _set(Zc, b, self._r2c(np.ascontiguousarray(y).view(np.float64), m, inner))

    def complex_to_real(self, sp, Zc, Xr):
        m, inner = sp.dims[0], sp.dims[1]
        for b in range(Zc.B):
            z = _get(Zc, b)
            if z.shape[0]:
                _set(Xr, b, self._c2r(z, m, inner).view(np.complex128))

    def hx_fused_real(self, sp, Ve, Vo, pot, fdiag, X, Y, fdiag_sep=None):
        """Y = H X on real-representation blocks: rotate to the complex basis, apply the complex operator
        (transform, potential, transform back, LC diagonal), project back."""
        from sqcircuit_b200.kernels import BlockView
        m, inner = sp.dims[0], sp.dims[1]
        nc = (inner - 1) // 2
        N = m * inner
        if inner % 2 == 0 or pot.n_terms > 4:
            return False
        V = np.zeros((m, m))
        V[0::2, :Ve.shape[1]] = Ve.numpy()
        V[1::2, :Vo.shape[1]] = Vo.numpy()
        # mirror columns of the parity-structured x basis: V[n][m-1-i] = (-1)^n V[n][i]
        sign = (-1.0) ** np.arange(m)
        for i in range(m // 2):
            V[:, m - 1 - i] = sign * V[:, i]
        Vt = torch.from_numpy(V)
        cmap = np.concatenate((nc - np.arange(nc, 0, -1), [nc], nc + np.arange(1, nc + 1)))   # complex col -> |n|
        cmap = np.abs(np.arange(inner) - nc)
        for b in range(X.B):
            y = _get(X, b)
            nv = y.shape[0]
            if nv == 0:
                continue
            z = torch.from_numpy(self._r2c(np.ascontiguousarray(y).view(np.float64), m, inner)).reshape(1, nv, N)
            s1, s2, out = torch.zeros_like(z), torch.zeros_like(z), torch.zeros_like(z)
            a, g, etab = pot._keep[0], pot._keep[1], pot._keep[2]

            class _P:
                pass
            pb = _P()
            pb.n_terms, pb.shift, pb.lam = pot.n_terms, pot.shift, pot.lam
            pb._keep = (a[b:b + 1], g[b:b + 1], etab[min(b, etab.shape[0] - 1)][None], None)
            self.mode_transform(sp, 0, Vt, 1, BlockView(z), BlockView(s1))
            self.potential_apply(sp, pb, BlockView(s1), BlockView(s2))
            self.mode_transform(sp, 0, Vt, 0, BlockView(s2), BlockView(out))
            r = self._c2r(out[0].numpy(), m, inner)
            if fdiag_sep is not None:
                sep = fdiag_sep.numpy()
                fd = np.zeros(2 * ((N + 1) // 2))
                fd[:N] = (sep[:m, None] + sep[None, m:]).reshape(-1)
                r = r + fd[None, :] * np.ascontiguousarray(y).view(np.float64)
            elif fdiag is not None:
                fd = fdiag[min(b, fdiag.shape[0] - 1)].numpy()
                r = r + fd[None, :] * np.ascontiguousarray(y).view(np.float64)
                if N % 2:
                    r[:, N:] = 0.0
            _set(Y, b, r.view(np.complex128))
        return True

    def block_update(self, A, Cf, Out, op=0, trans=0, impl=0):
        for b in range(A.B):
            nb = _cnt(Out, b)
            if nb == 0:
                continue
            a = _get(A, b)
            na = a.shape[0]
            if na == 0 and op == 1:
                continue
            cf = Cf[b].numpy()
            co = cf[:na, :nb].T if trans else cf[:nb, :na]
            r = co @ a
            if op == 1:
                r = _get(Out, b) - r
            _set(Out, b, r)

    def ritz_update(self, A, A2, Cf, Out, Out2, theta, rn2):
        for b in range(A.B):
            nb = _cnt(Out, b)
            if nb == 0:
                continue
            a, a2 = _get(A, b).copy(), _get(A2, b).copy()
            co = Cf[b].numpy()[:nb, :a.shape[0]]
            x, hx = co @ a, co @ a2
            _set(Out, b, x)
            _set(Out2, b, hx)
            rn2[b] = 0
            for j in range(nb):
                r = hx[j] - float(theta[b, j]) * x[j]
                rn2[b, j] = float(np.vdot(r, r).real)

    def block_rightmul(self, T, Cq, nout):
        for b in range(T.B):
            nin = _cnt(T, b)
            if nin == 0:
                continue
            t = _get(T, b)
            no = int(nout[b])
            r = np.zeros_like(t)
            r[:no] = Cq[b].numpy()[:no, :nin] @ t
            _set(T, b, r)

    def heev(self, A, n, w, Zt):
        B, ld, _ = A.shape
        for b in range(B):
            nb = n if isinstance(n, int) else int(n[b])
            if nb <= 0:
                continue
            a = A[b, :nb, :nb].numpy()
            ww, v = np.linalg.eigh((a + a.conj().T) / 2)
            w[b] = float('inf')
            w[b, :nb] = torch.from_numpy(ww)
            Zt[b, :nb, :nb] = torch.from_numpy(np.ascontiguousarray(v.T))

    def residual_norms(self, X, HX, theta, rn2):
        p = X.cnt_fixed
        for b in range(X.B):
            x, hx = _get(X, b), _get(HX, b)
            rn2[b] = 0
            for j in range(x.shape[0]):
                r = hx[j] - float(theta[b, j]) * x[j]
                rn2[b, j] = float(np.vdot(r, r).real)

    def davidson_state(self, **kw):
        class S:
            pass
        st = S()
        for k, v in kw.items():
            setattr(st, k, v)
        return st

    def download(self, t):
        return t.detach().cpu().numpy().copy()

    def upload(self, arr, dtype=None):
        t = torch.as_tensor(np.array(arr))
        return t.to(dtype) if dtype is not None else t

    def davidson_decide(self, st, B):
        st.n_not_done[0] = 0
        p, k = st.p, st.k
        for b in range(B):
            if int(st.done[b]):
                st.nact[b] = 0
                st.restart[b] = 0
                continue
            msz = int(st.msz[b])
            pe = min(p, msz)
            th = st.theta[b].numpy()
            want = th[:min(k, pe)]
            finite = bool(np.all(np.isfinite(want))) and pe >= min(k, p)
            scale = max(st.tol_abs_floor, np.max(np.abs(want[np.isfinite(want)]), initial=0.0))
            rmax, n = (0.0 if finite else 1e300), 0
            for j in range(pe):
                if not (np.isfinite(th[j]) and np.isfinite(float(st.rn2[b, j]))):
                    continue
                rel = np.sqrt(float(st.rn2[b, j])) / scale
                if j < k:
                    rmax = max(rmax, rel)
                if rel > 0.5 * st.tol_rel and j < (getattr(st, 'nexp', 0) or p):
                    st.act[b, n] = j
                    n += 1
            st.resmax[b] = rmax
            restart = 0
            planned = getattr(st, 'xoff', None) is not None
            if planned:
                st.xlast[b] = st.xoff[b]
            if rmax <= st.tol_rel:
                st.done[b] = 1
                n = 0
            else:
                st.n_not_done[0] += 1
                if planned:
                    restart = int(st.restart[b])
                    mcur = pe if restart else msz
                    if mcur + n > st.mmax:
                        n = st.mmax - mcur
                elif msz + n > st.mmax:
                    restart = 1
            st.nact[b] = n
            if not planned:
                st.restart[b] = restart
            elif not int(st.done[b]):
                mnext = (pe if restart else msz) + n
                rnext = int(mnext + n > st.mmax)
                st.restart[b] = rnext
                st.xoff[b] = 0 if rnext else st.xrow
            hi = pe - 1
            while hi > 0 and not np.isfinite(th[hi]):
                hi -= 1
            fl = max(abs(th[0]), abs(th[hi]), th[hi] - th[0]) if np.isfinite(th[0]) else 0.0
            st.den_floor[b] = max(1e-3 * fl, 1e-300)
            if restart:
                st.msz[b] = pe
                st.G[b] = 0
                for r in range(p):
                    st.G[b, r, r] = th[r]
                if getattr(st, 'GB', None) is not None:
                    st.GB[b] = 0
                    for r in range(p):
                        st.GB[b, r, r] = 1.0
            if getattr(st, 'mtot', None) is not None:
                st.mtot[b] = int(st.msz[b]) + n

    def restart_copy(self, X, HX, V, W, restart):
        for b in range(V.B):
            if int(restart[b]):
                _set(V, b, _get(X, b))
                _set(W, b, _get(HX, b))

    def precond_residual(self, X, HX, T, theta, act, d, den_floor, real=False):
        for b in range(X.B):
            n = _cnt(T, b)
            if n == 0:
                continue
            x, hx = _rows(X, b), _rows(HX, b)
            if real:
                x, hx = np.ascontiguousarray(x).view(np.float64), np.ascontiguousarray(hx).view(np.float64)
            dd = d[min(b, d.shape[0] - 1)].numpy()
            fl = float(den_floor[b])
            out = []
            for jj in range(n):
                a = int(act[b, jj])
                th = float(theta[b, a])
                den = dd - th
                den = np.where(np.abs(den) < fl, np.where(den < 0, -fl, fl), den)
                out.append((hx[a] - th * x[a]) / den)
            out = np.array(out)
            _set(T, b, out.view(np.complex128) if real else out)

    def lowblock_assemble(self, sp, S, Dt, pot, d, sigma, unit, out, real=False):
        if real:
            return self._lowblock_assemble_real(sp, S, Dt, pot, d, sigma, unit, out)
        a, g, etab, diag = pot._keep
        m, inner = sp.dims[0], sp.dims[1]
        Dtn = Dt.numpy()
        B, ns, _ = out.shape
        for b in range(B):
            s = S[min(b, S.shape[0] - 1)].numpy()
            dd = d[min(b, d.shape[0] - 1)].numpy()
            n, c = s // inner, s % inner
            A = np.zeros((ns, ns), complex)
            A[np.arange(ns), np.arange(ns)] = dd[s] - float(sigma[b])
            same_c = c[:, None] == c[None, :]
            g0, g1 = float(g[b, 0]), float(g[b, 1])
            A += same_c * (g0 * (n[:, None] == n[None, :]) +
                           g1 * (np.abs(n[:, None] - n[None, :]) == 1) * np.sqrt(np.maximum(n[:, None], n[None, :])))
            for t in range(pot.n_terms):
                sh = int(pot.shift[t][1])
                at = complex(a[b, t])
                A += (c[None, :] == c[:, None] - sh) * at * Dtn[t][n[:, None], n[None, :]]
                A += (c[None, :] == c[:, None] + sh) * np.conj(at * Dtn[t][n[None, :], n[:, None]])
            out[b] = torch.from_numpy((A / unit).astype(np.complex64))

    def lowblock_assemble_generic(self, sp, S, Dt, pot, d, sigma, unit, out):
        """Contract of sqc_lowblock_assemble_generic: H_SS - sigma for any layout from per-mode tables."""
        a, g, etab, diag = pot._keep
        M = sp.n_modes
        dims = [sp.dims[i] for i in range(M)]
        harm = [bool(sp.harmonic[i]) for i in range(M)]
        strides = np.concatenate((np.cumprod(dims[::-1])[::-1][1:], [1])).astype(int)
        dtoff, off = [], 0
        for i in range(M):
            dtoff.append(off)
            if harm[i] and dims[i] > 1:
                off += dims[i] * dims[i]
        B, ns, _ = out.shape
        for b in range(B):
            s = S[min(b, S.shape[0] - 1)].numpy().astype(int)
            dd = d[min(b, d.shape[0] - 1)].numpy()
            dt = Dt[min(b, Dt.shape[0] - 1)].numpy()
            idx = [(s // strides[i]) % dims[i] for i in range(M)]
            differs = np.stack([idx[i][:, None] != idx[i][None, :] for i in range(M)])       # [M][ns][ns]
            ndiff = differs.sum(axis=0)
            A = np.zeros((ns, ns), complex)
            A[np.arange(ns), np.arange(ns)] = dd[s] - float(sigma[b]) + float(g[b, 0])
            for i in range(M):
                if harm[i]:
                    n, n2 = idx[i][:, None], idx[i][None, :]
                    only = (ndiff == 1) & differs[i] & (np.abs(n - n2) == 1)
                    A += only * float(g[b, 1 + i]) * np.sqrt(np.maximum(n, n2))
            for t in range(pot.n_terms):
                fw = np.ones((ns, ns), bool)
                bw = np.ones((ns, ns), bool)
                pf = np.ones((ns, ns), complex)
                pb = np.ones((ns, ns), complex)
                for i in range(M):
                    n, n2 = idx[i][:, None], idx[i][None, :]
                    if harm[i]:
                        if dims[i] > 1:
                            Di = dt[t, dtoff[i]:dtoff[i] + dims[i] ** 2].reshape(dims[i], dims[i])
                            pf = pf * Di[n, n2]
                            pb = pb * Di[n2, n]
                    else:
                        w = int(pot.shift[t][i])
                        fw &= (n2 == n - w)
                        bw &= (n2 == n + w)
                at = complex(a[b, t])
                A += fw * at * pf + bw * np.conj(at * pb)
            out[b] = torch.from_numpy((A / unit).astype(np.complex64 if out.dtype == torch.complex64 else np.complex128))

    def hamiltonian_dense(self, sp, Dt, pot, d, out):
        """Contract of sqc_hamiltonian_dense: the whole H of every point, complex128, from per-mode tables."""
        B, N, _ = out.shape
        S = torch.arange(N, dtype=torch.int32).reshape(1, N)
        if out.dtype == torch.float64:
            tmp = torch.empty(out.shape, dtype=torch.complex128)
            self.lowblock_assemble_generic(sp, S, Dt, pot, d, torch.zeros(B, dtype=torch.float64), 1.0, tmp)
            out.copy_(tmp.real)
        else:
            self.lowblock_assemble_generic(sp, S, Dt, pot, d, torch.zeros(B, dtype=torch.float64), 1.0, out)

    def syev(self, A, w, Zt):
        """Contract of sqc_syev."""
        for b in range(A.shape[0]):
            a = A[b].numpy()
            ev, vec = np.linalg.eigh(0.5 * (a + a.T))
            w[b] = torch.from_numpy(ev)
            Zt[b] = torch.from_numpy(np.ascontiguousarray(vec.T))

    def _lowblock_assemble_real(self, sp, S, Dt, pot, d, sigma, unit, out):
        """H_r = U^H H_c U on the real index set (each real state is a combination of charges +-n)."""
        a, g = pot._keep[0], pot._keep[1]
        m, inner = sp.dims[0], sp.dims[1]
        nc = (inner - 1) // 2
        Dtn = Dt.numpy()
        B, ns, _ = out.shape
        for b in range(B):
            s = S[min(b, S.shape[0] - 1)].numpy()
            dd = d[min(b, d.shape[0] - 1)].numpy()
            # complex closure: every (Fock, +-|n|) that a member of the set touches
            comp, U = {}, []
            for r, idx in enumerate(s):
                i, j = idx // inner, idx % inner
                n = j if j <= nc else j - nc
                for q, coef in (((n, 1.0),) if j == 0 else
                                ((n, 2 ** -0.5), (-n, 2 ** -0.5)) if j <= nc else
                                ((n, 1j * 2 ** -0.5), (-n, -1j * 2 ** -0.5))):
                    key = (i, q)
                    if key not in comp:
                        comp[key] = len(comp)
                    U.append((comp[key], r, coef))
            keys = sorted(comp, key=comp.get)
            n_ = np.array([k[0] for k in keys])
            c = np.array([k[1] for k in keys])
            nk = len(keys)
            A = np.zeros((nk, nk), complex)
            same_c = c[:, None] == c[None, :]
            g0, g1 = float(g[b, 0]), float(g[b, 1])
            A += same_c * (g0 * (n_[:, None] == n_[None, :]) +
                           g1 * (np.abs(n_[:, None] - n_[None, :]) == 1) * np.sqrt(np.maximum(n_[:, None], n_[None, :])))
            for t in range(pot.n_terms):
                sh = int(pot.shift[t][1])
                at = complex(a[b, t])
                A += (c[None, :] == c[:, None] - sh) * at * Dtn[t][n_[:, None], n_[None, :]]
                A += (c[None, :] == c[:, None] + sh) * np.conj(at * Dtn[t][n_[None, :], n_[:, None]])
            Um = np.zeros((nk, ns), complex)
            for ci, r, coef in U:
                Um[ci, r] = coef
            Hr = (Um.conj().T @ A @ Um).real
            Hr[np.arange(ns), np.arange(ns)] += dd[s] - float(sigma[b])
            out[b] = torch.from_numpy((Hr / unit).astype(np.float32 if out.dtype == torch.float32 else np.complex64))

    def hpd_inverse_c64(self, A):
        for b in range(A.shape[0]):
            A[b] = torch.from_numpy(np.linalg.inv(A[b].numpy().astype(complex)).astype(np.complex64))

    def lowblock_apply(self, X, HX, T, theta, act, S, Ainv, unit, group=1, real=False):
        for b in range(X.B):
            n = _cnt(T, b)
            if n == 0:
                continue
            s = S[min(b // group, S.shape[0] - 1)].numpy()
            x, hx = _rows(X, b), _rows(HX, b)
            t = _get(T, b).copy()
            if real:
                x, hx, t = (np.ascontiguousarray(v).view(np.float64) for v in (x, hx, t))
            ai = Ainv[b // group].numpy().astype(complex)
            for jj in range(n):
                av = int(act[b, jj])
                r = hx[av][s] - float(theta[b, av]) * x[av][s]
                t[jj][s] = (ai @ r / unit).real if real else ai @ r / unit
            _set(T, b, t.view(np.complex128) if real else t)

    def lowband_factor(self, A, blk_off, nblk, nbmax):
        """block LDL^H of the block-tridiagonal matrix: diagonal blocks <- inverse Schur complements"""
        for b in range(A.shape[0]):
            a = A[b].numpy().astype(complex)
            off = blk_off[b].numpy()
            pprev = None
            for q in range(int(nblk[b])):
                o, o1 = int(off[q]), int(off[q + 1])
                m = a[o:o1, o:o1].copy()
                if q > 0:
                    op = int(off[q - 1])
                    e = a[op:o, o:o1]
                    m = m - e.conj().T @ pprev @ e
                pprev = np.linalg.inv(m)
                a[o:o1, o:o1] = pprev
            A[b] = torch.from_numpy(a.real.astype(np.float32) if A.dtype == torch.float32 else a.astype(np.complex64))

    def lowband_apply(self, X, HX, T, theta, act, S, Afac, blk_off, nblk, nbmax, unit, group=1, perm=None,
                      real=False):
        for b in range(X.B):
            n = _cnt(T, b)
            if n == 0:
                continue
            sl = b // group
            s = S[min(sl, S.shape[0] - 1)].numpy()
            a = Afac[sl].numpy().astype(complex)
            off = blk_off[sl].numpy()
            nb = int(nblk[sl])
            x, hx = _rows(X, b), _rows(HX, b)
            t = _get(T, b).copy()
            if real:
                x, hx, t = (np.ascontiguousarray(v).view(np.float64) for v in (x, hx, t))
            for jj in range(n):
                av = int(act[b, jj])
                v = ((hx[av][s] - float(theta[b, av]) * x[av][s]) / unit).astype(complex)
                for q in range(nb):
                    o, o1 = int(off[q]), int(off[q + 1])
                    y = v[o:o1].copy()
                    if q > 0:
                        op = int(off[q - 1])
                        y = y - a[op:o, o:o1].conj().T @ v[op:o]
                    v[o:o1] = a[o:o1, o:o1] @ y
                for q in range(nb - 2, -1, -1):
                    o, o1, o2 = int(off[q]), int(off[q + 1]), int(off[q + 2])
                    v[o:o1] = v[o:o1] - a[o:o1, o:o1] @ (a[o:o1, o1:o2] @ v[o1:o2])
                t[jj][s] = v.real if real else v
            _set(T, b, t.view(np.complex128) if real else t)

    def svqb_coeffs(self, S, nact, drop, Cq, msz=None, mtot=None):
        B, ld, _ = S.shape
        for b in range(B):
            n = min(int(nact[b]), ld)
            Cq[b] = 0
            if n <= 0:
                if mtot is not None:
                    mtot[b] = msz[b]
                continue
            s = S[b, :n, :n].numpy()
            dg = np.real(np.diag(s))
            dsc = np.where(dg > 0, 1 / np.sqrt(np.where(dg > 0, dg, 1)), 0.0)
            a = (s + s.conj().T) / 2 * np.outer(dsc, dsc)
            w, u = np.linalg.eigh(a)
            order = [i for i in np.argsort(-w) if w[i] > drop * w.max() and w[i] > 0]
            for j, src in enumerate(order):
                Cq[b, j, :n] = torch.from_numpy(dsc * u[:, src] / np.sqrt(w[src]))
            nact[b] = len(order)
            if mtot is not None:
                mtot[b] = int(msz[b]) + len(order)

    def davidson_insert(self, st, S, B):
        for b in range(B):
            na, msz = int(st.nact[b]), int(st.msz[b])
            if na <= 0:
                continue
            s = S[b].numpy()
            blk = s[:msz + na, :na].copy()
            d = blk[msz:, :]
            blk[msz:, :] = (d + d.conj().T) / 2
            G = st.G[b].numpy()
            G[:msz + na, msz:msz + na] = blk
            G[msz:msz + na, :msz + na] = blk.conj().T
            st.msz[b] = msz + na
