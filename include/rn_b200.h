/* rn_b200.h -- C ABI of librn_b200.so: the B200 (sm_100a) sweep-site kernels behind
 * renormalizer.mps.backend.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream handle (cudaStream_t
 * passed as void*); none takes or returns a torch / cupy type.  All functions return 0 on
 * success or a cudaError_t value.  Tensors are dense, row-major (C order); a complex128 tensor is
 * the usual interleaved (re, im) pair of doubles.  `cplx` = 0 -> float64, 1 -> complex128.
 *
 * Each function names the reference interface it replaces (path:line in shuaigroup/Renormalizer).
 */
#ifndef RN_B200_H
#define RN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* number of partial sums a reduction workspace must hold per vector: ws needs
 * 2 * RN_REDUCE_BLOCKS * nvec doubles */
#define RN_REDUCE_BLOCKS 296

/* Library / device bring-up; returns the SM count in *sm_count and the compute capability in
 * *cc (e.g. 100).  Fails (non-zero) when the device is not an sm_100 part. */
int rn_init(int device, int* sm_count, int* cc);
/* Version string of the library ("rn_b200 <n>"), for the loader's self check. */
const char* rn_version(void);

/* ---- contraction primitives ----------------------------------------------------------------
 * xp.tensordot (renormalizer/mps/matrix.py:210) lowered to one K-major real GEMM:
 *   C[i*ldc + j] (+)= sum_k A[i*lda + k] * B[j*ldb + k]
 * batch > 1 runs independent problems at the given element strides. */
int rn_dgemm_tn(void* stream, int m, int n, int k, const double* A, long lda, const double* B,
                long ldb, double* C, long ldc, int accumulate, int batch, long strideA,
                long strideB, long strideC);

/* out (M x N) = a (M x K) . b (K x N), all row-major, float64 or complex128: xp.tensordot with one
 * contracted axis / xp.dot (renormalizer/mps/matrix.py:210) as ONE call.  path as for rn_hop_plan:
 * 0 = FP64 DMMA, 1 = tcgen05 int8 split GEMM for products large enough to fill tiles. */
int rn_matmul(void* stream, int cplx, int M, int K, int N, const void* a, const void* b, void* out,
              int path);

/* Same contraction with FP64 accuracy on the tcgen05 tensor cores: both operands are split into
 * `nslices` (1..8) signed 8-bit digits per element (Ozaki scheme), the digit products run as
 * int8 tcgen05.mma with int32 TMEM accumulators and are recombined in FP64.  nslices = 7 bounds
 * the error by ~1e-14 * K * max|A_i| * max|B_j|; nslices = 8 reaches FP64 round-off. */
int rn_ozaki_gemm_tn(void* stream, int m, int n, int k, const double* A, long lda, const double* B,
                     long ldb, double* C, long ldc, int nslices);

/* Digits per element (1..8, default 7) used when a plan runs with path = 1, and the m*n*k below
 * which a contraction stays on the exact DMMA kernel (negative: keep the current value). */
int rn_set_ozaki(int nslices, double min_work);

/* Strided view -> K-major GEMM operand.  Element (r, c) of the source is src[r*s_row + c*s_col]
 * (strides in elements).  mode 0 ("A-form"): dst[r*dst_ld + c] (complex: interleaved, conj_flag
 * conjugates).  mode 1 ("B-form", complex only): 2x2 real representation, rows (2r, 2r+1), so a
 * complex product is a single real GEMM; conj_flag folds a conjugation of the LEFT operand in.
 * dst_ld is in doubles. */
int rn_pack(void* stream, int cplx, int mode, int conj_flag, int rows, int cols, const void* src,
            long s_row, long s_col, double* dst, long dst_ld);

/* MPO-site application  out[x,d,y1,f,y2] = sum_{p,q} W[p,d,q,f] in[x,p,q,y],  y = y1*Y2 + y2.
 * W in CSR over the output pair (d*F + f): rowptr[D*F+1], ent_pq[e] = p*Q + q, ent_val[e].
 * All strides in elements.  (Middle step of hop_expr.py:75-115 and lib.py:214-258.) */
int rn_wapply(void* stream, int cplx, const void* in, void* out, int X, int P, int Q, int Y,
              long isx, long isp, long isq, long isy, int D, int F, int Y2, long osx, long osd,
              long osf, long osy1, long osy2, const int* rowptr, const int* ent_pq,
              const double* ent_val);

/* ---- H_eff * C -----------------------------------------------------------------------------
 * Replaces renormalizer/mps/hop_expr.py:7 hop_expr(ltensor, rtensor, cmo, cshape) -> expr and the
 * returned expr(cstruct).  L is (La,Lb,Lc) = (bra, MPO, ket) bonds, R is (Rl,Rf,Rk); nsite = 0,
 * 1, 2 centre sites with physical dims d1,d2 and ancilla dims g1,g2 (1 when the state is an MPS,
 * >1 for an MPDM; hop_expr.py:80-91,104-115).  MPO site i is passed as CSR of W_i[p=left bond,
 * D=up, q=down, F=right bond].  L, R and the CSR arrays must stay alive while the plan lives.
 * path: 0 = FP64 DMMA GEMM, 1 = tcgen05 int8 split GEMM (FP64-accurate), see DESIGN.md. */
typedef struct rn_hop_plan rn_hop_plan;
int rn_hop_plan_create(rn_hop_plan** out, void* stream, int cplx, int nsite, const void* L,
                       int La, int Lb, int Lc, const void* R, int Rl, int Rf, int Rk, int d1,
                       int g1, int d2, int g2, int w1_F, const int* w1_rowptr, const int* w1_pq,
                       const double* w1_val, int w2_F, const int* w2_rowptr, const int* w2_pq,
                       const double* w2_val, int path);
/* out = H_eff . c_in ; c_in (Lc, d1[,g1][,d2[,g2]], Rk), out (La, ..., Rl). */
int rn_hop_apply(rn_hop_plan* plan, void* stream, const void* c_in, void* out);
/* kernels launched through this plan so far (for bench.py's gpu_launches) */
long rn_hop_plan_launches(const rn_hop_plan* plan);
int rn_hop_plan_destroy(rn_hop_plan* plan, void* stream);

/* ---- environment update --------------------------------------------------------------------
 * Replaces renormalizer/mps/lib.py:172 contract_one_site(environ, ms, mo, domain, ms_conj).
 * domain 0 = "L": env (Ea,Eb,Ec), bra (Ea,d,g,Mf), ket (Ec,d,g,Mh), W CSR as for hop.
 * domain 1 = "R": env (Ea,Eb,Ec), bra (Mf,d,g,Ea), ket (Mh,d,g,Ec), W CSR of
 *                 W'[p=right bond, D=up, q=down, F=left bond].
 * bra is the UN-conjugated bra-side site (== ket for an expectation value); out is (Mf,F,Mh). */
int rn_env_update(void* stream, int cplx, int domain, const void* env, int Ea, int Eb, int Ec,
                  const void* bra, const void* ket, int d, int g, int Mf, int Mh, int F,
                  const int* rowptr, const int* pq, const double* val, void* out, int path);

/* ---- bond decomposition --------------------------------------------------------------------
 * Replace the per-block scipy.linalg.qr / rq / svd calls of renormalizer/mps/svd_qn.py:170-186
 * (and optimized_svd, svd_qn.py:13-49).  A is (m x n) row-major with leading dimension lda (in
 * elements); k = min(m, n).
 *   rn_qr : A = Q R,  Q (m x k) orthonormal columns, R (k x n) upper trapezoidal (LAPACK signs)
 *   rn_lq : A = L Q,  L (m x k) lower trapezoidal,   Q (k x n) orthonormal rows
 *   rn_svd_jacobi : A = U diag(S) Vh, U (m x k), Vh (k x n); S is NOT sorted; *sweeps_out (host
 *                   int, may be NULL) receives the number of Jacobi sweeps, NEGATED when the last
 *                   sweep still rotated (no convergence within max_sweeps).  The bare iteration.
 *   rn_svd        : the same decomposition, QR-preconditioned (norm-ordered QR, QR of R^H, Jacobi
 *                   on the k x k factor): what scipy.linalg.svd(..., lapack_driver="gesdd") is
 *                   replaced by in optimized_svd (svd_qn.py:13-49) for every block that is not
 *                   tiny; `path` selects the GEMM path of the two back-multiplications. */
int rn_qr(void* stream, int cplx, int m, int n, const void* A, long lda, void* Q, long ldq,
          void* R, long ldr);
int rn_lq(void* stream, int cplx, int m, int n, const void* A, long lda, void* L, long ldl,
          void* Q, long ldq);
int rn_svd_jacobi(void* stream, int cplx, int m, int n, const void* A, long lda, void* U,
                  long ldu, double* S, void* Vh, long ldvh, int max_sweeps, int* sweeps_out);
int rn_svd(void* stream, int cplx, int m, int n, const void* A, long lda, void* U, long ldu,
           double* S, void* Vh, long ldvh, int max_sweeps, int path, int* sweeps_out);

/* ---- Krylov / Davidson vector kernels -------------------------------------------------------
 * (renormalizer/lib/krylov/krylov.py:55-83, renormalizer/lib/davidson/davidson.py:56-70,493-500)
 * n counts elements, nd counts doubles (2n for complex).  Scalars stay on the device.
 *   rn_multi_dot     out[2i..2i+1] = <V_i, x> = sum conj(V[i*ld+k]) x[k]   (ld in doubles)
 *   rn_lanczos_update w -= alpha[0]*vj + beta_prev[0]*vjm1 ; beta_out[0] = |w|   (vjm1 may be NULL)
 *   rn_scale_inv     out = x / s[0]
 *   rn_lincomb       out = sum_i coef[i] V_i          (coef complex when cplx) */
int rn_multi_dot(void* stream, int cplx, long n, int nvec, const double* V, long ld,
                 const double* x, double* ws, double* out);
int rn_lanczos_update(void* stream, long nd, double* w, const double* vj, const double* vjm1,
                      const double* alpha, const double* beta_prev, double* ws, double* beta_out);
int rn_scale_inv(void* stream, long nd, const double* x, const double* s, double* out);
int rn_lincomb(void* stream, int cplx, long n, int nvec, const double* V, long ld,
               const double* coef, double* out);
/* xp.allclose(a, b) of krylov.py:79: *violations (device int) = number of thread blocks that saw an
 * element with !(|a-b| <= atol + rtol |b|); 0 means close. */
int rn_allclose(void* stream, int cplx, long n, const double* a, const double* b, double rtol,
                double atol, int* violations);

/* One fused Lanczos iteration on the Krylov stack V (row j = v_j, n elements per row):
 *   w = H_eff v_j; alpha[j] = Re<v_j,w>; w -= alpha[j] v_j + beta[j-1] v_{j-1}; beta[j] = |w|;
 *   v_{j+1} = w / beta[j].   alpha / beta hold (value, 0) pairs on the device; V needs row j+1. */
int rn_lanczos_step(rn_hop_plan* plan, void* stream, long n, double* V, int j, double* alpha,
                    double* beta, double* w, double* ws);

/* The whole of expm_krylov(Afunc, dt, vstart) (renormalizer/lib/krylov/krylov.py:28-84) for an
 * H_eff plan: out = expm(dt * H_eff) v_in with dt = dt_re + i dt_im (dt_im must be 0 for real
 * vectors), same Lanczos recurrence and the same stopping rules as the reference (breakdown
 * 100 n eps, numpy.allclose of successive approximations every second step from the fifth on);
 * *nsteps_out = number of H_eff applications used.  The loop runs in C++, the tridiagonal
 * eigenproblem and the convergence test on the device; the host reads three integers per check.
 * Returns cudaErrorNotSupported when the Krylov dimension would exceed rn_krylov_max_dim()
 * (the caller then uses the step-wise entry points above). */
int rn_expm_krylov(rn_hop_plan* plan, void* stream, int cplx, long n, const void* v_in,
                   double dt_re, double dt_im, void* out, int* nsteps_out);
int rn_krylov_max_dim(void);

/* ---- Davidson eigensolver (DMRG local solver) -----------------------------------------------
 * Replaces renormalizer/lib/davidson/davidson.py:73 davidson(aop, x0, precond, ...) as called by
 * renormalizer/mps/gs.py:538-576 for nroots == 1: lowest eigenpair of  inverse * sum_p H_eff[p]
 * restricted to the entries with mask[k] != 0 (mask NULL: all), preconditioner 1/(hdiag - e + 1e-4)
 * (gs.py:512-514).  plans: nplans H_eff plans over the same centre tensor (the members of a stacked
 * MPO); x0: start vector (n elements, zero outside the mask); hdiag: n doubles.  c_out receives the
 * Ritz vector (not sign-fixed), *e_out the eigenvalue, *nhop_out the number of H_eff applications,
 * *converged_out 1 when |de| < tol and |r| < sqrt(tol) was reached within max_cycle.  The schedule
 * (subspace max_space, restart, lindep) is the reference's. */
int rn_davidson(rn_hop_plan** plans, int nplans, void* stream, int cplx, long n, const void* x0,
                const unsigned char* mask, const double* hdiag, double inverse, double tol,
                int max_cycle, int max_space, double lindep, void* c_out, double* e_out,
                int* nhop_out, int* converged_out);

/* ---- host-buffer entry points (what a NumPy-side caller binds) -------------------------------
 * Same contractions with HOST pointers: inputs are copied to the device, the kernels above run,
 * the result is copied back and the stream is synchronised.  W is the dense MPO site(s)
 * (Wb, d, d, Wf), real.  Shapes as for the device entry points. */
int rn_hop_apply_host(int cplx, int nsite, const void* L, int La, int Lb, int Lc, const void* R,
                      int Rl, int Rf, int Rk, int d1, int g1, int d2, int g2, const double* W1,
                      int w1_F, const double* W2, int w2_F, const void* c_in, void* out, int path);
int rn_env_update_host(int cplx, int domain, const void* env, int Ea, int Eb, int Ec,
                       const void* bra, const void* ket, int d, int g, int Mf, int Mh,
                       const double* W, int Wb, int Wf, void* out, int path);

/* scipy.linalg.svd(a, full_matrices=False) of optimized_svd (renormalizer/mps/svd_qn.py:13-49) with
 * HOST buffers: A (m x n) row-major in, U (m x k), S (k, descending), Vh (k x n) out. */
int rn_svd_host(int cplx, int m, int n, const void* A, void* U, double* S, void* Vh, int path);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * Between rn_profile_begin and rn_profile_end every contraction GEMM launch is bracketed by CUDA
 * events on its own stream; rn_profile_end returns the summed launch time, the summed
 * 2*m*n*k FLOPs and the launch count. */
/* Number of kernels this library has launched so far in this process. */
long rn_launch_count(void);
int rn_profile_begin(void);
int rn_profile_end(double* total_ms, double* total_flops, long* launches);
/* Dense int8 tensor-core rate of the device (the digit GEMM's roofline denominator): one CTA per SM
 * issues `iters` resident-operand 128x128x128 tcgen05.mma kind::i8 products; *tops_out in 1e12 op/s,
 * timed with CUDA events on `stream`. */
int rn_int8_peak(void* stream, int iters, double* tops_out);

#ifdef __cplusplus
}
#endif
#endif /* RN_B200_H */
