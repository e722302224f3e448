// One-sided (Hestenes) Jacobi SVD of a bond matrix block: A = U diag(S) V^H.
//
// Columns are held as contiguous rows (At[c*ldt + r] = A[r][c]) so the three inner products and
// the plane rotation of a column pair are coalesced streams reduced with warp shuffles.  A sweep
// is n-1 rounds of a round-robin tournament; the n/2 disjoint pairs of one round run in parallel,
// one block per pair.  Rotations are accumulated into V the same way.  High relative accuracy of
// the small singular values (Demmel-Veselic) is why Jacobi is used for the truncation step.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

namespace rn {

constexpr int J_THREADS = 256;

template <bool CPLX>
__global__ void __launch_bounds__(J_THREADS)
jacobi_round_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                    typename std::conditional<CPLX, double2, double>::type* __restrict__ Vw,
                    int m, int n, int nv, long ldt, long ldv, int round, int N, double tol,
                    int* __restrict__ rotated) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ double scratch[4 * 32];
  int p, q;
  if (blockIdx.x == 0) { p = N - 1; q = round; }
  else { p = (round + blockIdx.x) % (N - 1); q = (round - (int)blockIdx.x + (N - 1)) % (N - 1); }
  if (p >= n || q >= n) return;
  if (p > q) { const int t = p; p = q; q = t; }
  T* xp = At + (long)p * ldt;
  T* xq = At + (long)q * ldt;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // |xp|^2, |xq|^2, Re<xp,xq>, Im<xp,xq>
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) {
      const double2 a = xp[r], b = xq[r];
      acc[0] += a.x * a.x + a.y * a.y;
      acc[1] += b.x * b.x + b.y * b.y;
      acc[2] += a.x * b.x + a.y * b.y;
      acc[3] += a.x * b.y - a.y * b.x;
    } else {
      const double a = xp[r], b = xq[r];
      acc[0] += a * a; acc[1] += b * b; acc[2] += a * b;
    }
  }
  block_sum<4>(acc, scratch);
  const double a = acc[0], b = acc[1];
  const double gabs = sqrt(acc[2] * acc[2] + acc[3] * acc[3]);
  if (gabs == 0.0 || gabs <= tol * sqrt(a) * sqrt(b)) return;
  if (threadIdx.x == 0) *rotated = 1;
  const double phr = acc[2] / gabs, phi = acc[3] / gabs;  // phase = g / |g|
  const double zeta = (b - a) / (2.0 * gabs);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  // xp' = c xp - s conj(phase) xq ;  xq' = s phase xp + c xq
  for (int pass = 0; pass < 2; ++pass) {
    T* up = pass == 0 ? xp : Vw + (long)p * ldv;
    T* uq = pass == 0 ? xq : Vw + (long)q * ldv;
    const int len = pass == 0 ? m : nv;
    for (int r = threadIdx.x; r < len; r += blockDim.x) {
      if constexpr (CPLX) {
        const double2 u = up[r], v = uq[r];
        // conj(phase) * v
        const double cvx = phr * v.x + phi * v.y, cvy = phr * v.y - phi * v.x;
        // phase * u
        const double pux = phr * u.x - phi * u.y, puy = phr * u.y + phi * u.x;
        up[r] = make_double2(c * u.x - s * cvx, c * u.y - s * cvy);
        uq[r] = make_double2(s * pux + c * v.x, s * puy + c * v.y);
      } else {
        const double u = up[r], v = uq[r];
        up[r] = c * u - s * phr * v;
        uq[r] = s * phr * u + c * v;
      }
    }
  }
}

// S[c] = |At[c]|, At[c] /= S[c]
template <bool CPLX>
__global__ void __launch_bounds__(J_THREADS)
jacobi_finalize_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                       int m, long ldt, double* __restrict__ S) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ double scratch[32];
  T* x = At + (long)blockIdx.x * ldt;
  double ss[1] = {0.0};
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) { const double2 a = x[r]; ss[0] += a.x * a.x + a.y * a.y; }
    else { const double a = x[r]; ss[0] += a * a; }
  }
  block_sum<1>(ss, scratch);
  const double sig = sqrt(ss[0]);
  if (threadIdx.x == 0) S[blockIdx.x] = sig;
  const double inv = sig > 0.0 ? 1.0 / sig : 0.0;
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) { double2 a = x[r]; a.x *= inv; a.y *= inv; x[r] = a; }
    else x[r] *= inv;
  }
}

template <bool CPLX>
__global__ void jacobi_eye_kernel(typename std::conditional<CPLX, double2, double>::type* Vw, int n, long ldv) {
  pdl_wait();
  const long total = (long)n * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / n), r = (int)(i % n);
    if constexpr (CPLX) Vw[(long)c * ldv + r] = make_double2(r == c ? 1.0 : 0.0, 0.0);
    else Vw[(long)c * ldv + r] = r == c ? 1.0 : 0.0;
  }
}

// Economic SVD of A (m x n row-major, lda): U (m x k, ldu), S (k), Vh (k x n, ldvh), k = min(m,n).
// Singular values are NOT sorted.  Returns the number of sweeps in *sweeps_out (host int).
template <bool CPLX>
static int svd_driver(cudaStream_t st, int m, int n, const void* A, long lda, void* U, long ldu,
                      double* S, void* Vh, long ldvh, int max_sweeps, int* sweeps_out) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const int es = CPLX ? 2 : 1;
  const bool wide = m < n;
  const int mt = wide ? n : m;   // tall problem: mt x nt
  const int nt = wide ? m : n;
  const long ldt = mt, ldv = nt;
  T *At = nullptr, *Vw = nullptr;
  int* flag = nullptr;
  RN_CHECK(cudaMallocAsync((void**)&At, sizeof(T) * (size_t)nt * ldt, st));
  RN_CHECK(cudaMallocAsync((void**)&Vw, sizeof(T) * (size_t)nt * ldv, st));
  RN_CHECK(cudaMallocAsync((void**)&flag, sizeof(int), st));
  int err;
  if (!wide) err = launch_pack(st, CPLX, 0, 0, n, m, A, 1, lda, (double*)At, ldt * es);   // At = A^T
  else err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, m, n, A, lda, 1, (double*)At, ldt * es);  // At = conj(A)
  if (err) return err;
  int nbe = (int)ceil_div((long)nt * nt, 256);
  if (nbe > 1184) nbe = 1184;
  { RN_LAUNCH(jacobi_eye_kernel<CPLX>, nbe, 256, 0, st, Vw, nt, ldv); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  const int N = (nt + 1) & ~1;  // even number of players
  const double tol = sqrt((double)mt) * 2.220446049250313e-16;
  int sweeps = 0;
  if (nt > 1) {
    for (; sweeps < max_sweeps; ++sweeps) {
      RN_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
      for (int round = 0; round < N - 1; ++round)
        { RN_LAUNCH(jacobi_round_kernel<CPLX>, N / 2, J_THREADS, 0, st, At, Vw, mt, nt, nt, ldt, ldv, round, N, tol, flag); rn::g_launches++; }
      RN_LAUNCH_CHECK();
      int h = 0;
      RN_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      RN_CHECK(cudaStreamSynchronize(st));
      if (!h) { ++sweeps; break; }
    }
  }
  if (sweeps_out) *sweeps_out = sweeps;
  { RN_LAUNCH(jacobi_finalize_kernel<CPLX>, nt, J_THREADS, 0, st, At, mt, ldt, S); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  if (!wide) {
    // U[r][c] = At[c][r];  Vh[c][j] = conj(V[j][c]) = conj(Vw[c][j])
    err = launch_pack(st, CPLX, 0, 0, m, nt, At, 1, ldt, (double*)U, ldu * es);
    if (err) return err;
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, nt, n, Vw, ldv, 1, (double*)Vh, ldvh * es);
  } else {
    // A = V' S U'^H:  U[j][c] = Vw[c][j];  Vh[c][r] = conj(U'[r][c]) = conj(At[c][r])
    err = launch_pack(st, CPLX, 0, 0, m, nt, Vw, 1, ldv, (double*)U, ldu * es);
    if (err) return err;
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, nt, n, At, ldt, 1, (double*)Vh, ldvh * es);
  }
  if (err) return err;
  RN_CHECK(cudaFreeAsync(At, st));
  RN_CHECK(cudaFreeAsync(Vw, st));
  RN_CHECK(cudaFreeAsync(flag, st));
  return 0;
}

}  // namespace rn

extern "C" int rn_svd_jacobi(void* stream, int cplx, int m, int n, const void* A, long lda,
                             void* U, long ldu, double* S, void* Vh, long ldvh, int max_sweeps,
                             int* sweeps_out) {
  if (m <= 0 || n <= 0) return 0;
  if (max_sweeps <= 0) max_sweeps = 40;
  return cplx ? rn::svd_driver<true>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                     max_sweeps, sweeps_out)
              : rn::svd_driver<false>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                      max_sweeps, sweeps_out);
}
