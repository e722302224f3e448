// Householder QR / LQ of a bond matrix (the canonicalisation step of every TDVP-PS site and of
// _push_cano).  Same reflector convention as LAPACK's geqrf/zlarfg, so R's diagonal is real and Q
// matches scipy.linalg.qr up to round-off on full-rank input; rank-deficient input still yields an
// orthonormal Q (unlike Cholesky-QR), which TDVP with padded bond dimensions relies on.
//
// The matrix is held "column-as-row": At[c*ldt + r] = A[r][c], so that every reflector and every
// trailing column is a contiguous, coalesced stream.  The factorisation itself is the blocked
// cluster-panel algorithm of qr_panel.cu; this file holds the driver (pack in, factor, extract R,
// pack Q out), and a launch-per-reflector fallback for shapes whose panel does not fit on chip:
// each block re-derives the reflector from the (read-only during that step) pivot column and
// updates its own group of trailing columns, so no grid-wide barrier is needed.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"
#include "qr_common.cuh"

#include <stdlib.h>

namespace rn {

constexpr int Q_THREADS = 1024;  // passes over a column are L2-latency bound: few rows per thread
constexpr int Q_CPB = 4;  // trailing columns per block

// Reflector of column j (LAPACK zlarfg): H = I - tau v v^H, v[j] = 1, H^H x = beta e_j.
template <bool CPLX>
__device__ __forceinline__ void make_reflector(const typename Cx<CPLX>::T* colj, int m, int j,
                                               double* scratch, typename Cx<CPLX>::T& tau,
                                               typename Cx<CPLX>::T& scale, double& beta) {
  using C = Cx<CPLX>;
  double ss[1] = {0.0};
  for (int r = j + 1 + threadIdx.x; r < m; r += blockDim.x) ss[0] += C::abs2(colj[r]);
  block_sum<1>(ss, scratch);
  const typename C::T alpha = colj[j];
  const double ar = C::re(alpha), ai = C::im(alpha);
  if (ss[0] == 0.0 && ai == 0.0) {
    tau = C::zero();
    scale = C::zero();
    beta = ar;
    return;
  }
  const double nrm = sqrt(ar * ar + ai * ai + ss[0]);
  beta = ar >= 0.0 ? -nrm : nrm;
  tau = C::make((beta - ar) / beta, -ai / beta);
  // scale = 1 / (alpha - beta)
  const double dr = ar - beta, di = ai;
  const double den = dr * dr + di * di;
  scale = C::make(dr / den, -di / den);
}

// One elimination step: At[c] <- H_j^H At[c] for the trailing columns c > j.
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_step_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, long ldt, int j,
                  typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                  double* __restrict__ rdiag) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_CPB * 32];
  const T* colj = At + (long)j * ldt;
  T tau, scale;
  double beta;
  make_reflector<CPLX>(colj, m, j, scratch, tau, scale, beta);
  if (blockIdx.x == 0) {
    T* vj = V + (long)j * ldt;
    for (int r = threadIdx.x; r < m; r += blockDim.x)
      vj[r] = r < j ? C::zero() : (r == j ? C::one() : C::mul(colj[r], scale));
    if (threadIdx.x == 0) { tau_out[j] = tau; rdiag[j] = beta; }
  }
  const int c0 = j + 1 + blockIdx.x * Q_CPB;
  if (c0 >= n) return;
  double acc[2 * Q_CPB];
#pragma unroll
  for (int i = 0; i < 2 * Q_CPB; ++i) acc[i] = 0.0;
  for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
    const T v = r == j ? C::one() : C::mul(colj[r], scale);
#pragma unroll
    for (int i = 0; i < Q_CPB; ++i) {
      if (c0 + i < n) {
        const T d = C::cmul(v, At[(long)(c0 + i) * ldt + r]);
        acc[2 * i] += C::re(d);
        acc[2 * i + 1] += C::im(d);
      }
    }
  }
  block_sum<2 * Q_CPB>(acc, scratch);
  const T ctau = C::conj(tau);
  for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
    const T v = r == j ? C::one() : C::mul(colj[r], scale);
#pragma unroll
    for (int i = 0; i < Q_CPB; ++i) {
      if (c0 + i < n) {
        const T w = C::mul(ctau, C::make(acc[2 * i], acc[2 * i + 1]));
        T* p = At + (long)(c0 + i) * ldt + r;
        *p = C::sub(*p, C::mul(v, w));
      }
    }
  }
}

// Qt[c] <- H_j Qt[c] for c in [j, k)
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_applyq_kernel(typename Cx<CPLX>::T* __restrict__ Qt, int m, int k, long ldt, int j,
                    const typename Cx<CPLX>::T* __restrict__ V,
                    const typename Cx<CPLX>::T* __restrict__ tau_in) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_CPB * 32];
  const T* vj = V + (long)j * ldt;
  const T tau = tau_in[j];
  const int c0 = j + blockIdx.x * Q_CPB;
  if (c0 >= k) return;
  double acc[2 * Q_CPB];
#pragma unroll
  for (int i = 0; i < 2 * Q_CPB; ++i) acc[i] = 0.0;
  for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
    const T v = vj[r];
#pragma unroll
    for (int i = 0; i < Q_CPB; ++i) {
      if (c0 + i < k) {
        const T d = C::cmul(v, Qt[(long)(c0 + i) * ldt + r]);
        acc[2 * i] += C::re(d);
        acc[2 * i + 1] += C::im(d);
      }
    }
  }
  block_sum<2 * Q_CPB>(acc, scratch);
  for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
    const T v = vj[r];
#pragma unroll
    for (int i = 0; i < Q_CPB; ++i) {
      if (c0 + i < k) {
        const T w = C::mul(tau, C::make(acc[2 * i], acc[2 * i + 1]));
        T* p = Qt + (long)(c0 + i) * ldt + r;
        *p = C::sub(*p, C::mul(v, w));
      }
    }
  }
}


template <bool CPLX>
__global__ void set_identity_rows_kernel(typename Cx<CPLX>::T* Qt, int m, int k, long ldt) {
  pdl_wait();
  using C = Cx<CPLX>;
  const long total = (long)k * m;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / m), r = (int)(i % m);
    Qt[(long)c * ldt + r] = (r == c) ? C::one() : C::zero();
  }
}

// Rout[j*ldr + c] = R[j][c] (upper trapezoid, k x n); with lq != 0 write the conjugate transpose
// instead: Lout[c*ldr + j] = conj(R[j][c]) (n x k lower trapezoid).
template <bool CPLX>
__global__ void extract_r_kernel(const typename Cx<CPLX>::T* __restrict__ At, const double* __restrict__ rdiag,
                                 int n, int k, long ldt, typename Cx<CPLX>::T* __restrict__ Rout,
                                 long ldr, int lq) {
  pdl_wait();
  using C = Cx<CPLX>;
  const long total = (long)k * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i / n), c = (int)(i % n);
    typename C::T v = c < j ? C::zero() : (c == j ? C::make(rdiag[j], 0.0) : At[(long)c * ldt + j]);
    if (lq) Rout[(long)c * ldr + j] = C::conj(v);
    else Rout[(long)j * ldr + c] = v;
  }
}

template <bool CPLX>
int qr_colmajor_panel(cudaStream_t st, int m, int n, typename Cx<CPLX>::T* At, long ldt,
                      typename Cx<CPLX>::T* V, typename Cx<CPLX>::T* tau, double* rdiag,
                      typename Cx<CPLX>::T* Qt);

static int g_qr_panel = -1;      // blocked cluster-panel QR (qr_panel.cu); RN_QR_PANEL=0 disables it

// QR of the m x n matrix held column-as-row in At.  The blocked cluster-panel factorisation of
// qr_panel.cu handles every shape whose panel slice fits on chip; beyond that (or with
// RN_QR_PANEL=0) the simple launch-per-reflector kernels above are the fallback.
template <bool CPLX>
static int qr_colmajor(cudaStream_t st, int m, int n, typename Cx<CPLX>::T* At, long ldt,
                       typename Cx<CPLX>::T* V, typename Cx<CPLX>::T* tau, double* rdiag,
                       typename Cx<CPLX>::T* Qt) {
  const int k = m < n ? m : n;
  if (g_qr_panel < 0) {
    const char* e = getenv("RN_QR_PANEL");
    g_qr_panel = e ? atoi(e) : 1;
  }
  if (g_qr_panel) {
    const int perr = qr_colmajor_panel<CPLX>(st, m, n, At, ldt, V, tau, rdiag, Qt);
    if (perr != 1) return perr;       // 1 = shape outside the cluster kernel's range: fall through
  }
  for (int j = 0; j < k; ++j) {
    int nb = (int)ceil_div(n - j - 1, Q_CPB);
    if (nb < 1) nb = 1;
    { RN_LAUNCH(house_step_kernel<CPLX>, nb, Q_THREADS, 0, st, At, m, n, ldt, j, V, tau, rdiag); rn::g_launches++; }
  }
  RN_LAUNCH_CHECK();
  int nbi = (int)ceil_div((long)k * m, 256);
  if (nbi > 1184) nbi = 1184;
  { RN_LAUNCH(set_identity_rows_kernel<CPLX>, nbi, 256, 0, st, Qt, m, k, ldt); rn::g_launches++; }
  for (int j = k - 1; j >= 0; --j) {
    const int nb = (int)ceil_div(k - j, Q_CPB);
    { RN_LAUNCH(house_applyq_kernel<CPLX>, nb, Q_THREADS, 0, st, Qt, m, k, ldt, j, V, tau); rn::g_launches++; }
  }
  RN_LAUNCH_CHECK();
  return 0;
}

// A (m x n row-major, lda)  ->  Q (m x k, ldq) R (k x n, ldr)          [lq == 0]
// A (m x n row-major, lda)  ->  L (m x k, ldr) Q (k x n, ldq)          [lq == 1], via QR of A^H
template <bool CPLX>
static int qr_driver(cudaStream_t st, int lq, int m, int n, const void* A, long lda, void* Q,
                     long ldq, void* R, long ldr) {
  using T = typename Cx<CPLX>::T;
  const int es = CPLX ? 2 : 1;  // doubles per element
  // tall problem dims: (mt x nt), columns stored as rows of At (nt x mt)
  const int mt = lq ? n : m, nt = lq ? m : n;
  const int k = mt < nt ? mt : nt;
  const long ldt = mt;
  T *At = nullptr, *V = nullptr, *Qt = nullptr, *tau = nullptr;
  double* rdiag = nullptr;
  RN_CHECK(cudaMallocAsync((void**)&At, sizeof(T) * (size_t)nt * ldt, st));
  RN_CHECK(cudaMallocAsync((void**)&V, sizeof(T) * (size_t)k * ldt, st));
  RN_CHECK(cudaMallocAsync((void**)&Qt, sizeof(T) * (size_t)k * ldt, st));
  RN_CHECK(cudaMallocAsync((void**)&tau, sizeof(T) * (size_t)k, st));
  RN_CHECK(cudaMallocAsync((void**)&rdiag, sizeof(double) * (size_t)k, st));
  int err;
  if (!lq) {
    // At[c][r] = A[r][c]: rows of the packed matrix are A's columns
    err = launch_pack(st, CPLX, 0, 0, n, m, A, 1, lda, (double*)At, ldt * es);
  } else {
    // tall matrix is A^H (n x m); its column c is conj(A[c][:]) -> At = conj(A)
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, m, n, A, lda, 1, (double*)At, ldt * es);
  }
  if (err) return err;
  err = qr_colmajor<CPLX>(st, mt, nt, At, ldt, V, tau, rdiag, Qt);
  if (err) return err;
  int nbr = (int)ceil_div((long)k * nt, 256);
  if (nbr > 1184) nbr = 1184;
  { RN_LAUNCH(extract_r_kernel<CPLX>, nbr, 256, 0, st, At, rdiag, nt, k, ldt, (T*)R, ldr, lq); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  if (!lq) {
    // Q[r][c] = Qt[c][r]
    err = launch_pack(st, CPLX, 0, 0, m, k, Qt, 1, ldt, (double*)Q, ldq * es);
  } else {
    // Q = (Qhat)^H: Q[c][r] = conj(Qhat[r][c]) = conj(Qt[c][r])
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, k, n, Qt, ldt, 1, (double*)Q, ldq * es);
  }
  if (err) return err;
  RN_CHECK(cudaFreeAsync(At, st));
  RN_CHECK(cudaFreeAsync(V, st));
  RN_CHECK(cudaFreeAsync(Qt, st));
  RN_CHECK(cudaFreeAsync(tau, st));
  RN_CHECK(cudaFreeAsync(rdiag, st));
  return 0;
}

}  // namespace rn

extern "C" int rn_qr(void* stream, int cplx, int m, int n, const void* A, long lda, void* Q,
                     long ldq, void* R, long ldr) {
  if (m <= 0 || n <= 0) return 0;
  return cplx ? rn::qr_driver<true>((cudaStream_t)stream, 0, m, n, A, lda, Q, ldq, R, ldr)
              : rn::qr_driver<false>((cudaStream_t)stream, 0, m, n, A, lda, Q, ldq, R, ldr);
}

extern "C" int rn_lq(void* stream, int cplx, int m, int n, const void* A, long lda, void* L,
                     long ldl, void* Q, long ldq) {
  if (m <= 0 || n <= 0) return 0;
  return cplx ? rn::qr_driver<true>((cudaStream_t)stream, 1, m, n, A, lda, Q, ldq, L, ldl)
              : rn::qr_driver<false>((cudaStream_t)stream, 1, m, n, A, lda, Q, ldq, L, ldl);
}
