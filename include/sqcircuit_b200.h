/*
 * sqcircuit_b200 — C ABI of the B200-native diagonalisation hot path.
 *
 * The reference (stanfordLINQS/SQcircuit) is pure Python; its hot path is
 *   Circuit._build_op_memory / _build_exp_ops        circuit.py:1621-1714   (operators)
 *   Circuit._get_LC_hamil / _get_inductive_hamil /
 *   Circuit.hamiltonian                              circuit.py:1731-1829, 2147-2164
 *   Circuit._diag_np  (scipy.sparse.linalg.eigs)     circuit.py:1938-1986
 *   Circuit.matrix_elements / dec_rate               circuit.py:2397-2647
 *   Sweep.sweepFlux / sweepCharge                    sweep.py:71-148
 * Each entry point below names the reference code it replaces.  A maintainer
 * binds them with ctypes; no torch types cross this ABI:
 * plain device pointers, sizes and a cudaStream_t passed as void*.
 *
 * Conventions
 *   - complex128 everywhere, interleaved (re, im) doubles  == torch.complex128.
 *   - a "block" is a set of state vectors per sweep point:  data[b][v][n],
 *     n fastest (n < N = prod(mode dims), reference Kronecker order: mode 0
 *     slowest), v < count(b), b < B.  See sqc_block_t.
 *   - small matrices per sweep point are row-major with a fixed leading dim.
 *   - every function returns 0 on success, a cudaError_t (>0) on a CUDA
 *     failure or a negative SQC_E* code on invalid arguments.
 */
#ifndef SQCIRCUIT_B200_H
#define SQCIRCUIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQC_MAX_MODES 8
#define SQC_MAX_TERMS 16      /* junction-like terms (distinct rows of W) */
#define SQC_EINVAL   (-1)
#define SQC_ELIMIT   (-2)

/* A batch of state-vector blocks living in one allocation. */
typedef struct {
    void*          ptr;        /* device, complex128 */
    int64_t        stride_b;   /* complex elements between consecutive sweep points */
    const int32_t* off;        /* device[B] first vector index per point, or NULL (0) */
    const int32_t* cnt;        /* device[B] number of vectors per point, or NULL (cnt_fixed) */
    int32_t        cnt_fixed;  /* count when cnt == NULL; always the maximum count */
} sqc_block_t;

/* Tensor-product space of the truncated modes (circuit.py:1340-1379). */
typedef struct {
    int32_t n_modes;
    int32_t dims[SQC_MAX_MODES];      /* m_i (charge modes: 2*trunc-1) */
    int32_t harmonic[SQC_MAX_MODES];  /* 1: Fock basis, 0: charge basis */
} sqc_space_t;

/* Potential part of a Hermitian circuit operator, expressed in the basis where
 * every harmonic mode is rotated to the eigenbasis of its truncated (a + a^dag)
 * ("x basis"; displacement operators D(alpha), alpha imaginary, are diagonal
 * there) and charge modes stay in the charge basis:
 *
 *   (P z)(i) = [ g0(b) + sum_m g_m(b) lam_m[i_m] + diag(i) ] z(i)
 *            + sum_t  a_t(b) ph_t(i) z(i - sh_t) + conj(a_t(b) ph_t(i)) z(i + sh_t)
 *   ph_t(i)  = prod_{m harmonic} etab[t][m][i_m],      sh_t = integer shift on charge modes
 *
 * This is sum_j -EJ_j cos(phi_j) plus the external-flux inductor terms of
 * Circuit._get_inductive_hamil (circuit.py:1788-1829) for a_t = -EJ e^{i phi_ext}/2,
 * and the cos/sin/sin_half operators of the noise code for other a_t. */
typedef struct {
    int32_t        n_terms;
    int32_t        shift[SQC_MAX_TERMS][SQC_MAX_MODES]; /* per-mode index shift (charge modes) */
    const double*  lam;        /* device[sum dims]  x-basis eigenvalues per mode (0 for charge) */
    const void*    etab;       /* device complex [(B or 1)][n_terms][sum dims] */
    int64_t        etab_stride_b; /* 0 when shared by all sweep points */
    const double*  g;          /* device[B][1 + n_modes]  linear coefficients */
    const void*    a;          /* device complex [B][n_terms] */
    const double*  diag;       /* optional device[(B or 1)][N] extra real diagonal, or NULL */
    int64_t        diag_stride_b;
} sqc_potential_t;

/* ---- library ---------------------------------------------------------- */
int         sqc_abi_version(void);   /* 4 */
const char* sqc_error_string(int code);
/* number of kernels launched by this library since load (bench.py gpu_launches) */
int64_t     sqc_launch_count(void);
/* SHA-256 (hex) of the kernel sources + this header the library was built from; the Python loader
 * refuses a library whose hash differs from the sources next to it (no stale binaries). */
const char* sqc_source_hash(void);

/* ---- (1) operator construction ---------------------------------------- */
/* x-basis of one harmonic mode: eigen-decomposition of the truncated a + a^dag
 * (real symmetric tridiagonal, m x m).  Replaces qt.displace(N, alpha).expm()
 * (circuit.py:1707-1708): D(i s) = V diag(exp(i s lam)) V^T.
 * V is row-major V[n][i] = <Fock n | x_i>, lam ascending.  m <= 112. */
int sqc_xbasis(int32_t m, double* V, double* lam, void* stream);

/* Phase tables etab[t][off_m + i] = exp(i * s[t][m] * lam_m[i]) for harmonic modes
 * (1 for charge modes).  s = phi_zp_m * wTrans[t][m]  (circuit.py:1603-1619).
 * s: device[(B or 1)][n_terms][n_modes]. */
int sqc_phase_tables(const sqc_space_t* sp, int32_t n_terms, const double* lam,
                     const double* s, int64_t s_stride_b, void* etab,
                     int64_t etab_stride_b, int32_t B, void* stream);

/* Diagonal of the LC Hamiltonian (circuit.py:1731-1757) in the Fock/charge basis:
 *   d(i) = sum_{m harm} omega_m i_m + sum_{m<=m' charge} q[m][m'] (n_m - ng_m)(n_m' - ng_m')
 * omega: device[(B or 1)][n_modes], q: device[(B or 1)][n_modes][n_modes] (upper triangle,
 * diagonal already halved as in _get_quadratic_cInvTrans, charge unit (2e)^2/hbar folded in),
 * ng: device[(B or 1)][n_modes].  out: device[(B or 1)][N]. */
int sqc_lc_diagonal(const sqc_space_t* sp, const double* omega, const double* q,
                    const double* ng, int64_t param_stride_b, double* out,
                    int32_t B_out, void* stream);

/* ---- (2) matrix-free operator apply ------------------------------------ */
/* Y = (accumulate ? Y : 0) + diag .* X */
int sqc_diag_mul(sqc_block_t X, sqc_block_t Y, const double* diag, int64_t diag_stride_b,
                 int64_t N, int32_t B, int32_t accumulate, void* stream);

/* Mode-wise contraction  Z[.., i, ..] = sum_k Mat[i][k] X[.., k, ..]  (transpose: Mat[k][i])
 * along mode `mode` of every vector in the block; optional epilogue
 * Z += ediag .* EX (used to add the LC diagonal on the way back to the Fock basis).
 * Mat: device[(B or 1)][m][m] real.  In-place (Z == X) is allowed.
 * impl: 0 = FP64 tensor-core (DMMA) kernel, 1 = plain FP64 SIMT kernel. */
int sqc_mode_transform(const sqc_space_t* sp, int32_t mode, const double* Mat,
                       int64_t mat_stride_b, int32_t transpose, sqc_block_t X, sqc_block_t Z,
                       const double* ediag, int64_t ediag_stride_b, sqc_block_t EX,
                       int32_t B, int32_t impl, void* stream);

/* Fused Y = fdiag .* X + V [ P ( V^T X ) ] in one kernel for spaces with exactly one harmonic
 * mode (mode 0) and one charge mode (mode 1) -- the 0-pi layout.  Uses the parity of the x
 * basis: Ve[a][i] = V[2a][i] (ceil(m/2) square), Vo[a][i] = V[2a+1][i] (floor(m/2) square),
 * i over the first half of the x points.  pot->diag must be NULL, shifts in {-1,0,1},
 * n_terms <= 4.  fdiag may be NULL.  Returns SQC_ELIMIT when the layout is not supported
 * (callers then use sqc_mode_transform + sqc_potential_apply).  Replaces the CSR mat-vec on
 * the matrix of Circuit.hamiltonian() (circuit.py:2147-2164). */
int sqc_hx_fused(const sqc_space_t* sp, const double* Ve, const double* Vo,
                 const sqc_potential_t* pot, const double* fdiag, int64_t fdiag_stride_b,
                 sqc_block_t X, sqc_block_t Y, int32_t B, void* stream);

/* Y = fdiag .* X + (x_i V_i) [ P ( (x_i V_i^T) X ) ] in ONE kernel for any layout whose harmonic modes (more than
 * one level each, at most four of them, at most 32 levels each) lead the Kronecker order, followed by up to eight
 * charge modes with junction shifts in {-1, 0, 1}: the operator sqc_mode_transform (forward, one call per
 * harmonic mode) + sqc_potential_apply + sqc_mode_transform (backward) compute in 2 n_harmonic + 1 passes over
 * the block.  Vcat: the x-basis matrices V_i (m_i x m_i, row-major, as returned by sqc_xbasis) of the harmonic
 * modes with m_i > 1, concatenated in mode order.  pot->diag must be NULL; fdiag [(B or 1)][N] may be NULL.
 * A CTA stages all harmonic indices x a window of charge columns (+ halo) in shared memory with TMA bulk copies.
 * Returns SQC_ELIMIT when the layout is not supported (callers then use the unfused sequence).
 * Replaces the CSR mat-vec on the matrix of Circuit.hamiltonian() (circuit.py:2147-2164). */
int sqc_hx_generic(const sqc_space_t* sp, const double* Vcat, const sqc_potential_t* pot, const double* fdiag,
                   int64_t fdiag_stride_b, sqc_block_t X, sqc_block_t Y, int32_t B, void* stream);

/* ---- real (time-reversal adapted) representation ------------------------------------------
 * For one harmonic mode (mode 0, dim m) x one charge mode (mode 1, dim inner = 2 nc + 1) at zero gate
 * charge the Hamiltonian commutes with T = (complex conjugation) x (n -> -n), so its eigenvectors can be
 * chosen with z(i, -n) = conj z(i, n).  Such a vector is stored as N = m * inner REAL numbers
 *     y[i][0] = z(i, 0),  y[i][n] = sqrt2 Re z(i, n),  y[i][nc + n] = sqrt2 Im z(i, n),  1 <= n <= nc,
 * an orthonormal basis in which H is real symmetric.  Blocks keep the sqc_block_t convention with
 * Nc = ceil(N / 2) complex elements per vector (two real amplitudes per element; the padding double of
 * an odd N is zero), so the Gram / update kernels of section (3) serve both representations: the real
 * part of a complex Gram entry of two such blocks is the real inner product.
 *
 * sqc_hx_fused_real: Y = fdiag .* X + V [ P ( V^T X ) ] on real vectors -- same operator, arguments and
 * limits as sqc_hx_fused, except: fdiag is the LC diagonal in the real layout, [(B or 1)][2 Nc]
 * (fdiag_stride_b in doubles), and inner must be odd and <= 127.  Vectors are staged with TMA bulk copies
 * (cp.async.bulk + mbarrier) and written back the same way.
 * fdiag_sep (optional, device [m + inner], shared by all sweep points): the same diagonal in separable form,
 * d(i, j) = fdiag_sep[i] + fdiag_sep[m + j] (LC energies of one harmonic and one charge mode add up); when
 * given it is used INSTEAD of fdiag: the charge part is applied in the x basis (exact, V is orthogonal) and
 * the epilogue only needs one Fock-energy scalar per row. */
int sqc_hx_fused_real(const sqc_space_t* sp, const double* Ve, const double* Vo,
                      const sqc_potential_t* pot, const double* fdiag, int64_t fdiag_stride_b,
                      const double* fdiag_sep, sqc_block_t X, sqc_block_t Y, int32_t B, void* stream);

/* Zc (complex, N elements per vector, Fock x charge basis of the reference) <- Xr (real representation),
 * and the projection back, Xr <- T-symmetric part of Zc. */
int sqc_real_to_complex(const sqc_space_t* sp, sqc_block_t Xr, sqc_block_t Zc, int32_t B, void* stream);
int sqc_complex_to_real(const sqc_space_t* sp, sqc_block_t Zc, sqc_block_t Xr, int32_t B, void* stream);

/* Z = P X with P as documented at sqc_potential_t (X, Z in the x/charge basis). */
int sqc_potential_apply(const sqc_space_t* sp, const sqc_potential_t* pot, sqc_block_t X,
                        sqc_block_t Z, int32_t B, void* stream);

/* out[b][i][j] = <z_i| P |z_j> for the first nv <= 4 vectors of every point (x/charge basis, P as documented at
 * sqc_potential_t): one read of the vectors, no vector written.  The matrix elements of the potential-type
 * operators of the noise code (cos / sin / sin_half of a junction, inductive coupling, dH/d phi_ext;
 * circuit.py:2397-2647) between two eigenstates rotated to the x basis once.  out: device complex [B][nv][nv];
 * work: device scratch of sqc_potential_elements_work_bytes() bytes (per-chunk partial sums, reduced in a fixed
 * order).  Z.cnt must be NULL. */
int64_t sqc_potential_elements_work_bytes(int64_t N, int32_t nv, int32_t B);
int sqc_potential_elements(const sqc_space_t* sp, const sqc_potential_t* pot, sqc_block_t Z, int32_t nv, void* out,
                           void* work, int32_t B, void* stream);

/* Y = (accumulate ? Y : 0) + c_dn(b) * (a_m X) + c_up(b) * (a_m^dag X) on harmonic mode m
 * (charge / flux operators of harmonic modes, circuit.py:1471-1524).  c_*: device complex[B]. */
int sqc_ladder_apply(const sqc_space_t* sp, int32_t mode, const void* c_dn, const void* c_up,
                     sqc_block_t X, sqc_block_t Y, int32_t B, int32_t accumulate, void* stream);

/* ---- (3) batched block-Davidson building blocks ------------------------- */
/* C[b][i][j] = sum_n conj(A[b][i][n]) Bm[b][j][n]    (Rayleigh-Ritz projections, Gram
 * matrices; FP64 tensor cores).  C: device complex [B][ldr][ldc] row-major, rows >= count(A)
 * and columns >= count(Bm) are written as zero up to (A.cnt_fixed, Bm.cnt_fixed).
 * work: device scratch of sqc_gram_work_bytes() bytes (split-N partial sums), or NULL when
 * that returns 0.  impl: 0 = DMMA, 1 = SIMT reference. */
int64_t sqc_gram_work_bytes(int32_t B, int32_t na_max, int32_t nb_max, int64_t N);
int sqc_gram(sqc_block_t A, sqc_block_t Bm, int64_t N, int32_t B, void* C, int32_t ldr,
             int32_t ldc, void* work, int32_t impl, void* stream);

/* One pass over A for two right-hand blocks: C = A^H Bm, C2 = A^H Bm2 (same leading dims;
 * work sized by sqc_gram_work_bytes).  Used for the new columns of the basis Gram matrix and of
 * the projected Hamiltonian in the non-orthogonalised Davidson iteration. */
int sqc_gram_pair(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int32_t B, void* C,
                  void* C2, int32_t ldr, int32_t ldc, void* work, void* stream);
/* sqc_gram_pair on blocks of the real representation (real arrays viewed as complex blocks, see
 * sqc_hx_fused_real): C = Re(A^H Bm), C2 = Re(A^H Bm2), imaginary parts zero.  A column tile of the tensor-core
 * product holds eight right-hand vectors instead of four (re, im) pairs: half the DMMAs. */
int sqc_gram_pair_real(sqc_block_t A, sqc_block_t Bm, sqc_block_t Bm2, int64_t N, int32_t B, void* C,
                       void* C2, int32_t ldr, int32_t ldc, void* work, void* stream);

/* Out[b][j][n] (op)= sum_i Cf[b][j][i] A[b][i][n],  i < count(A), j < count(Out);
 * op: 0 assign, 1 subtract from Out.  Cf: device complex [B][ldr][ldc]; the coefficient of
 * A_i in Out_j is Cf[j][i] (trans = 0) or Cf[i][j] (trans = 1).
 * impl: 0 = FP64 tensor-core kernel, 1 = plain FP64 SIMT kernel. */
int sqc_block_update(sqc_block_t A, const void* Cf, int32_t ldr, int32_t ldc, int32_t trans,
                     sqc_block_t Out, int64_t N, int32_t B, int32_t op, int32_t impl, void* stream);

/* In-place T[b][j][n] <- sum_i Cq[b][j][i] T[b][i][n], i < count(T) (<= 16), j < nout[b]. */
int sqc_block_rightmul(sqc_block_t T, const void* Cq, int32_t ldr, int32_t ldc,
                       const int32_t* nout, int64_t N, int32_t B, void* stream);

/* Batched Hermitian eigensolver (parallel cyclic Jacobi, one CTA per matrix).
 * A: device complex [B][ld][ld] (read only; full Hermitian storage), n: device[B] sizes or
 * NULL (n_fixed).  w: device[B][ld] ascending eigenvalues (entries >= n set to +inf),
 * Zt: device complex [B][ld][ld], Zt[j][i] = component i of eigenvector j.  ld <= 112. */
int sqc_heev(const void* A, const int32_t* n, int32_t n_fixed, int32_t ld, double* w, void* Zt,
             int32_t B, void* stream);

/* Generalised Rayleigh-Ritz GA c = theta GB c (GB = Gram matrix of a non-orthogonal basis):
 * Cholesky of GB (vectors whose pivot falls below drop * GB_ii are removed), Jacobi on
 * L^-1 GA L^-H, c = L^-H y.  w ascending, Zt[j][i] = coefficient i of Ritz vector j. ld <= 68. */
int sqc_gen_rr(const void* GA, const void* GB, const int32_t* n, int32_t ld, double drop, double* w,
               void* Zt, int32_t B, void* stream);

/* sqc_gen_rr for a REAL symmetric pencil (Davidson basis of the real representation): only the real parts
 * of GA / GB are read, the arithmetic is real, and only the n_out lowest Ritz vectors are written to Zt
 * (rows [0, min(n_out, n)), zero imaginary parts; the other rows are left untouched).  w as sqc_gen_rr.
 * One 384-thread CTA per sweep point, five resident per SM (mmax = 40). */
int sqc_gen_rr_real(const void* GA, const void* GB, const int32_t* n, int32_t ld, double drop, double* w,
                    void* Zt, int32_t n_out, int32_t B, void* stream);

/* Ritz block in one pass over the basis A and its image A2 (same counts):
 *   Out[b][j] = sum_i Cf[b][j][i] A[b][i],  Out2[b][j] = sum_i Cf[b][j][i] A2[b][i],
 *   rn2[b][j] = || Out2_j - theta[b][j] Out_j ||^2,   j < count(Out, b)   (row stride Out.cnt_fixed).
 * Equivalent to two sqc_block_update(op = 0, trans = 0) calls plus sqc_residual_norms without reading
 * the Ritz vectors back.  Out may alias the leading rows of A (and Out2 of A2): every position reads all
 * its inputs before it is written.  Sweep points with count(Out, b) == 0 are left untouched (rn2 too).
 * work: sqc_ritz_update_work_bytes(B, N) bytes of device scratch. */
int64_t sqc_ritz_update_work_bytes(int32_t B, int64_t N);
int sqc_ritz_update(sqc_block_t A, sqc_block_t A2, const void* Cf, int32_t ldr, int32_t ldc,
                    sqc_block_t Out, sqc_block_t Out2, const double* theta, int32_t ld_theta,
                    double* rn2, void* work, int64_t N, int32_t B, void* stream);

/* rn2[b][j] = || HX_j - theta[b][j] X_j ||^2,  j < p. */
int sqc_residual_norms(sqc_block_t X, sqc_block_t HX, const double* theta, int32_t ld_theta,
                       int64_t N, int32_t B, double* rn2, void* stream);

/* Davidson bookkeeping for one iteration (one small CTA per sweep point); see
 * sqcircuit_b200/eigensolver.py for the algorithm.  All arrays device, per point. */
typedef struct {
    int32_t  p, k, mmax, ldg;   /* block size, wanted pairs, max basis, leading dim of G/Cf */
    double   tol_rel;           /* residual tolerance relative to max(|theta_0..k-1|, floor_abs) */
    double   tol_abs_floor;
    int32_t* msz;               /* [B] basis size */
    int32_t* nact;              /* [B] number of new (active) vectors */
    int32_t* act;               /* [B][p] Ritz index of each active vector */
    int32_t* restart;           /* [B] restart flag for this iteration */
    int32_t* done;              /* [B] converged flag */
    int32_t* n_not_done;        /* [1] counter (zeroed by the callee) */
    double*  theta;             /* [B][ldg] Ritz values */
    double*  rn2;               /* [B][p] squared residual norms */
    double*  resmax;            /* [B] max relative residual over the k wanted pairs */
    double*  den_floor;         /* [B] preconditioner denominator floor */
    void*    G;                 /* [B][ldg][ldg] projected Hamiltonian */
    void*    GB;                /* [B][ldg][ldg] Gram matrix of the basis (non-orthogonalised
                                   variant), or NULL */
    int32_t* mtot;              /* [B] msz + nact after sqc_davidson_decide, or NULL */
    int32_t  nexp;              /* only Ritz pairs j < nexp are expanded with a correction vector
                                   (0 = all p; k = the wanted pairs only, the guard vectors ride along) */
    /* Copy-free restarts (both NULL: the caller copies X, HX into the basis with sqc_restart_copy).
     * The basis buffers have xrow + p rows; the Ritz block of an iteration is written either to
     * rows [xrow, xrow + p) or, when this iteration restarts, straight into rows [0, p).  The
     * restart of iteration i+1 is decided at iteration i (basis after the insert plus as many new
     * vectors again would exceed mmax), so the row is known before the block is computed:
     *   xoff[b]  (in/out) row for the NEXT Ritz block, restart[b] (in/out) the matching flag;
     *   xlast[b] (out)    row of the Ritz block this call has judged (frozen once done[b]).
     * The new-vector count is clipped to the free rows when no restart was scheduled. */
    int32_t* xoff;
    int32_t* xlast;
    int32_t  xrow;
    /* 1: the basis vectors are REAL arrays viewed as complex blocks (real representation, see
     * sqc_hx_fused_real): only the real parts of the Gram / projection entries are meaningful and
     * sqc_davidson_insert2 stores them with a zero imaginary part. */
    int32_t  real_basis;
} sqc_davidson_state_t;

int sqc_davidson_decide(const sqc_davidson_state_t* st, int32_t B, void* stream);

/* restart: where restart[b], V[b][0..p) <- X[b], W[b][0..p) <- HX[b]. */
int sqc_restart_copy(sqc_block_t X, sqc_block_t HX, sqc_block_t V, sqc_block_t W,
                     const int32_t* restart, int64_t N, int32_t B, void* stream);

/* T[b][jj] = (HX_a - theta_a X_a) / clamp(d - theta_a),  a = act[b][jj], jj < nact[b];
 * T addresses the free tail of the basis (T.off = msz). d: device[(B or 1)][N]. */
int sqc_precond_residual(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                         int32_t ld_theta, const int32_t* act, int32_t ld_act,
                         const double* d, int64_t d_stride_b, const double* den_floor,
                         int64_t N, int32_t B, void* stream);

/* sqc_precond_residual on a real-representation block: d has one entry per REAL amplitude,
 * device[(B or 1)][2 N] (N = complex elements per vector, d_stride_b in doubles, even). */
int sqc_precond_residual_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                              int32_t ld_theta, const int32_t* act, int32_t ld_act,
                              const double* d, int64_t d_stride_b, const double* den_floor,
                              int64_t N, int32_t B, void* stream);

/* Two-level preconditioner (harmonic x charge layouts): the ns basis states S[.] with the smallest
 * LC energy get the exact inverse of (H_SS - sigma) instead of the diagonal one.
 *  - assemble: out[b] (complex64 [ns][ns]) = (H_SS(b) - sigma[b] I) / unit, built analytically from
 *    the Fock-basis displacement tables Dt[t] = V diag(etab_t) V^T (complex128 [n_terms][m][m]);
 *  - inverse: in-place Gauss-Jordan in shared memory (ns <= 160);
 *  - apply: T[b][jj][S[r]] = sum_q Ainv[b][r][q] (HX_a - theta_a X_a)[S[q]] / unit, a = act[b][jj]
 *    (overwrites the diagonal-preconditioned entries written by sqc_precond_residual).
 *    group >= 1: sweep points b share the index set and inverse of slot b / group (S: [ceil(B/group)]
 *    rows, Ainv: [ceil(B/group)][ns][ns]) -- neighbouring points of a sweep have nearly the same low
 *    block, and a preconditioner only has to be approximate. */
int sqc_lowblock_assemble(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                          const void* Dt, const sqc_potential_t* pot, const double* d,
                          int64_t d_stride_b, const double* sigma, double unit, void* out,
                          int32_t B, void* stream);

/* The same block for ANY layout (multi-mode sweeps, batches of circuits): Dt holds per harmonic mode i with
 * m_i > 1 the Fock-basis matrix of every junction term on that mode, V_i diag(etab_t^(i)) V_i^T, concatenated in mode
 * order per term: device complex [(B or 1)][n_terms][sum_i m_i^2] (dt_stride_b = 0 when shared).  S: linear indices
 * of the ns kept basis states, [(B or 1)][ns]. */
int sqc_lowblock_assemble_generic(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                                  const void* Dt, int64_t dt_stride_b, const sqc_potential_t* pot, const double* d,
                                  int64_t d_stride_b, const double* sigma, double unit, void* out, int32_t B,
                                  void* stream);
/* The whole Hamiltonian of every sweep point as a dense matrix, out[b][r][q] = <r|H(b)|q> (complex128,
 * [B][N][N], N = prod dims <= 4096), assembled from the same per-mode tables as above -- what the reference
 * materialises in Circuit.hamiltonian() (circuit.py:2147-2164) before a dense / sparse eigensolve.  Feeds
 * sqc_heev (or, real_out, sqc_syev) on the dense path (N <= 112). */
int sqc_hamiltonian_dense(const sqc_space_t* sp, const void* Dt, int64_t dt_stride_b, const sqc_potential_t* pot,
                          const double* d, int64_t d_stride_b, void* out, int32_t real_out, int32_t B, void* stream);
/* real_out != 0: out is double [B][N][N] (layouts without charge modes only: there a_t D_t + h.c. = 2 Re(a_t D_t),
 * H is real symmetric), the input of sqc_syev: batched real symmetric eigensolver (parallel cyclic Jacobi, matrix
 * and eigenvectors in shared memory, one CTA per matrix).  A: device double [B][ld][ld] (read only), n <= 112 the
 * size of every matrix, w: device [B][ld] ascending eigenvalues (entries >= n set to +inf), Zt: device double
 * [B][ld][ld], Zt[j][i] = component i of eigenvector j.  Replaces the eigs / eigh call of Circuit._diag_np
 * (circuit.py:1938-1986) for small real Hamiltonians. */
int sqc_syev(const double* A, int32_t n, int32_t ld, double* w, double* Zt, int32_t B, void* stream);
int sqc_hpd_inverse_c64(void* A, int32_t ns, int32_t B, void* stream);
int sqc_lowblock_apply(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                       int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                       int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                       int64_t N, int32_t B, void* stream);

/* Real representation: S holds indices of REAL basis states (i * inner + j), d the real-layout LC
 * diagonal; the block is real symmetric.  real_storage = 0: stored as complex64 with zero imaginary parts
 * (for sqc_hpd_inverse_c64 + sqc_lowblock_apply_real, the dense variant); real_storage = 1: stored as
 * float [ns][ns] (for sqc_lowband_factor_real + sqc_lowband_apply_real).  The *_apply_real variants gather
 * and scatter real amplitudes (N = complex elements per vector, as everywhere). */
int sqc_lowblock_assemble_real(const sqc_space_t* sp, int32_t ns, const int32_t* S, int64_t s_stride_b,
                               const void* Dt, const sqc_potential_t* pot, const double* d,
                               int64_t d_stride_b, const double* sigma, double unit, void* out,
                               int32_t real_storage, int32_t B, void* stream);
int sqc_lowblock_apply_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                            int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                            int64_t s_stride_b, const void* Ainv, int32_t ns, double unit, int32_t group,
                            int64_t N, int32_t B, void* stream);
/* Banded variant in float storage: same contract as sqc_lowband_factor / sqc_lowband_apply below with a
 * real symmetric block; a warp carries four right-hand sides as one float4 per row. */
int sqc_lowband_factor_real(void* A, int32_t ns, const int32_t* blk_off, int32_t ld_off, const int32_t* nblk,
                            int32_t nbmax, int32_t B, void* stream);
int sqc_lowband_apply_real(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                           int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                           int64_t s_stride_b, const void* Afac, int32_t ns, const int32_t* blk_off,
                           int32_t ld_off, const int32_t* nblk, int32_t nbmax, double unit, int32_t group,
                           const int32_t* perm, int64_t N, int32_t B, void* stream);

/* Banded form of the two-level preconditioner for LARGE low-energy blocks (ns up to 1024).
 * With the index set ordered charge-major (all kept Fock states of charge q form block q, index
 * range [blk_off[q], blk_off[q+1])), H_SS is block tridiagonal (junction charge shifts in {-1,0,1}).
 *  - factor: A ([B][ns][ns] complex64 from sqc_lowblock_assemble with that ordering) gets its diagonal
 *    blocks overwritten by the inverses P_q of the block-LDL^H Schur complements; blocks (q, q+1) stay.
 *    blk_off: [B][ld_off] int32, nblk: [B], nbmax: largest block (<= 96).
 *  - apply: same contract as sqc_lowblock_apply, via forward / backward block substitution
 *    (cost ~ 3 sum_q n_q^2 per right-hand side instead of ns^2).  perm (same shape and stride as S, or
 *    NULL): a permutation that sorts each index set by address; gather and scatter then walk memory
 *    in ascending order (coalesced runs) instead of the charge-major block order. */
int sqc_lowband_factor(void* A, int32_t ns, const int32_t* blk_off, int32_t ld_off, const int32_t* nblk,
                       int32_t nbmax, int32_t B, void* stream);
int sqc_lowband_apply(sqc_block_t X, sqc_block_t HX, sqc_block_t T, const double* theta,
                      int32_t ld_theta, const int32_t* act, int32_t ld_act, const int32_t* S,
                      int64_t s_stride_b, const void* Afac, int32_t ns, const int32_t* blk_off,
                      int32_t ld_off, const int32_t* nblk, int32_t nbmax, double unit, int32_t group,
                      const int32_t* perm, int64_t N, int32_t B, void* stream);

/* SVQB coefficients from the Gram matrix S = T^H T of the new block:
 * Cq[j][i] = D_i U_ij sigma_j^-1/2 over directions with sigma_j > drop * sigma_max;
 * nact[b] <- number kept; mtot[b] <- msz[b] + nact[b] when both pointers are given.
 * S: [B][ld][ld] (read only), Cq: [B][ld][ld], ld <= 16. */
int sqc_svqb_coeffs(const void* S, int32_t ld, int32_t* nact, double drop, void* Cq,
                    const int32_t* msz, int32_t* mtot, int32_t B, void* stream);

/* G[b][i][msz+j] = S[b][i][j], G[b][msz+j][i] = conj(.), i < msz+nact, j < nact; msz += nact. */
int sqc_davidson_insert(const sqc_davidson_state_t* st, const void* S, int32_t ldr, int32_t ldc,
                        int32_t B, void* stream);

/* as sqc_davidson_insert, for both G (from SA) and GB (from SB). */
int sqc_davidson_insert2(const sqc_davidson_state_t* st, const void* SA, const void* SB, int32_t ldr,
                         int32_t ldc, int32_t B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SQCIRCUIT_B200_H */
