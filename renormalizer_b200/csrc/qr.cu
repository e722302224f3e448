// Householder QR / LQ of a bond matrix (the canonicalisation step of every TDVP-PS site and of
// _push_cano).  Same reflector convention as LAPACK's geqrf/zlarfg, so R's diagonal is real and Q
// matches scipy.linalg.qr up to round-off on full-rank input; rank-deficient input still yields an
// orthonormal Q (unlike Cholesky-QR), which TDVP with padded bond dimensions relies on.
//
// The matrix is held "column-as-row": At[c*ldt + r] = A[r][c], so that every reflector and every
// trailing column is a contiguous, coalesced stream.  One launch per reflector; each block
// re-derives the reflector from the (read-only during that step) pivot column and updates its own
// group of trailing columns, so no grid-wide barrier is needed.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace rn {

constexpr int Q_THREADS = 256;
constexpr int Q_CPB = 4;  // trailing columns per block

template <bool CPLX>
struct Cx;
template <>
struct Cx<false> {
  using T = double;
  __device__ static T zero() { return 0.0; }
  __device__ static T one() { return 1.0; }
  __device__ static T mul(T a, T b) { return a * b; }
  __device__ static T cmul(T a, T b) { return a * b; }  // conj(a) * b
  __device__ static T sub(T a, T b) { return a - b; }
  __device__ static T conj(T a) { return a; }
  __device__ static double abs2(T a) { return a * a; }
  __device__ static double re(T a) { return a; }
  __device__ static double im(T) { return 0.0; }
  __device__ static T make(double r, double) { return r; }
};
template <>
struct Cx<true> {
  using T = double2;
  __device__ static T zero() { return make_double2(0.0, 0.0); }
  __device__ static T one() { return make_double2(1.0, 0.0); }
  __device__ static T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  __device__ static T cmul(T a, T b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
  __device__ static T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
  __device__ static T conj(T a) { return make_double2(a.x, -a.y); }
  __device__ static double abs2(T a) { return a.x * a.x + a.y * a.y; }
  __device__ static double re(T a) { return a.x; }
  __device__ static double im(T a) { return a.y; }
  __device__ static T make(double r, double i) { return make_double2(r, i); }
};

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


// ---- single-launch variants ------------------------------------------------------------------
// Elimination of all k reflectors in ONE cooperative launch: column c is owned by block
// c % gridDim.x for the whole factorisation, a grid barrier separates the steps (column j+1 must
// be final before every block derives reflector j+1 from it).
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_factor_coop_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, int k, long ldt,
                         typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                         double* __restrict__ rdiag) {
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_CPB * 32];
  cg::grid_group grid = cg::this_grid();
  const int nb = gridDim.x, bid = blockIdx.x;
  for (int j = 0; j < k; ++j) {
    const T* colj = At + (long)j * ldt;
    T tau, scale;
    double beta;
    make_reflector<CPLX>(colj, m, j, scratch, tau, scale, beta);
    if (bid == (j % nb)) {
      T* vj = V + (long)j * ldt;
      for (int r = threadIdx.x; r < m; r += blockDim.x)
        vj[r] = r < j ? C::zero() : (r == j ? C::one() : C::mul(colj[r], scale));
      if (threadIdx.x == 0) { tau_out[j] = tau; rdiag[j] = beta; }
    }
    const T ctau = C::conj(tau);
    // owned columns c > j, c == bid (mod nb), Q_CPB at a time
    int c = j + 1 + ((bid - (j + 1)) % nb + nb) % nb;
    while (c < n) {
      int cols[Q_CPB];
#pragma unroll
      for (int i = 0; i < Q_CPB; ++i) { cols[i] = c < n ? c : -1; c += nb; }
      double acc[2 * Q_CPB];
#pragma unroll
      for (int i = 0; i < 2 * Q_CPB; ++i) acc[i] = 0.0;
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : C::mul(colj[r], scale);
#pragma unroll
        for (int i = 0; i < Q_CPB; ++i)
          if (cols[i] >= 0) {
            const T d = C::cmul(v, At[(long)cols[i] * ldt + r]);
            acc[2 * i] += C::re(d);
            acc[2 * i + 1] += C::im(d);
          }
      }
      block_sum<2 * Q_CPB>(acc, scratch);
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : C::mul(colj[r], scale);
#pragma unroll
        for (int i = 0; i < Q_CPB; ++i)
          if (cols[i] >= 0) {
            const T w = C::mul(ctau, C::make(acc[2 * i], acc[2 * i + 1]));
            T* p = At + (long)cols[i] * ldt + r;
            *p = C::sub(*p, C::mul(v, w));
          }
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// Q = H_0 ... H_{k-1} I in ONE launch: each block owns Q_CPB columns of Q and applies the
// reflectors j = c_max .. 0 to them (H_j leaves e_c untouched for j > c).
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_formq_kernel(typename Cx<CPLX>::T* __restrict__ Qt, int m, int k, long ldt,
                   const typename Cx<CPLX>::T* __restrict__ V,
                   const typename Cx<CPLX>::T* __restrict__ tau_in) {
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_CPB * 32];
  const int c0 = blockIdx.x * Q_CPB;
  if (c0 >= k) return;
  const int ncol = (k - c0) < Q_CPB ? (k - c0) : Q_CPB;
  for (int i = 0; i < ncol; ++i)
    for (int r = threadIdx.x; r < m; r += blockDim.x)
      Qt[(long)(c0 + i) * ldt + r] = (r == c0 + i) ? C::one() : C::zero();
  __syncthreads();
  for (int j = c0 + ncol - 1; j >= 0; --j) {
    const T* vj = V + (long)j * ldt;
    const T tau = tau_in[j];
    double acc[2 * Q_CPB];
#pragma unroll
    for (int i = 0; i < 2 * Q_CPB; ++i) acc[i] = 0.0;
    for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
      const T v = vj[r];
#pragma unroll
      for (int i = 0; i < Q_CPB; ++i)
        if (i < ncol && c0 + i >= j) {
          const T d = C::cmul(v, Qt[(long)(c0 + i) * ldt + r]);
          acc[2 * i] += C::re(d);
          acc[2 * i + 1] += C::im(d);
        }
    }
    block_sum<2 * Q_CPB>(acc, scratch);
    for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
      const T v = vj[r];
#pragma unroll
      for (int i = 0; i < Q_CPB; ++i)
        if (i < ncol && c0 + i >= j) {
          const T w = C::mul(tau, C::make(acc[2 * i], acc[2 * i + 1]));
          T* p = Qt + (long)(c0 + i) * ldt + r;
          *p = C::sub(*p, C::mul(v, w));
        }
    }
    __syncthreads();
  }
}

template <bool CPLX>
__global__ void set_identity_rows_kernel(typename Cx<CPLX>::T* Qt, int m, int k, long ldt) {
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
  using C = Cx<CPLX>;
  const long total = (long)k * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i / n), c = (int)(i % n);
    typename C::T v = c < j ? C::zero() : (c == j ? C::make(rdiag[j], 0.0) : At[(long)c * ldt + j]);
    if (lq) Rout[(long)c * ldr + j] = C::conj(v);
    else Rout[(long)j * ldr + c] = v;
  }
}

static int g_qr_coop_blocks[2] = {-1, -1};   // co-resident block budget per dtype, -1 = unknown

template <bool CPLX>
static int qr_colmajor(cudaStream_t st, int m, int n, typename Cx<CPLX>::T* At, long ldt,
                       typename Cx<CPLX>::T* V, typename Cx<CPLX>::T* tau, double* rdiag,
                       typename Cx<CPLX>::T* Qt) {
  const int k = m < n ? m : n;
  int& budget = g_qr_coop_blocks[CPLX ? 1 : 0];
  if (budget < 0) {
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    RN_CHECK(cudaGetDevice(&dev));
    RN_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    RN_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, house_factor_coop_kernel<CPLX>,
                                                           Q_THREADS, 0));
    budget = coop ? sms * (per_sm > 2 ? 2 : per_sm) : 0;
  }
  if (budget > 0) {
    int nb = n < budget ? n : budget;
    if (nb < 1) nb = 1;
    void* args[] = {(void*)&At, (void*)&m, (void*)&n, (void*)&k, (void*)&ldt, (void*)&V, (void*)&tau,
                    (void*)&rdiag};
    RN_CHECK(cudaLaunchCooperativeKernel((void*)house_factor_coop_kernel<CPLX>, dim3(nb), dim3(Q_THREADS),
                                         args, 0, st));
    house_formq_kernel<CPLX><<<(unsigned)ceil_div(k, Q_CPB), Q_THREADS, 0, st>>>(Qt, m, k, ldt, V, tau);
    RN_LAUNCH_CHECK();
    return 0;
  }
  for (int j = 0; j < k; ++j) {
    int nb = (int)ceil_div(n - j - 1, Q_CPB);
    if (nb < 1) nb = 1;
    house_step_kernel<CPLX><<<nb, Q_THREADS, 0, st>>>(At, m, n, ldt, j, V, tau, rdiag);
  }
  RN_LAUNCH_CHECK();
  int nbi = (int)ceil_div((long)k * m, 256);
  if (nbi > 1184) nbi = 1184;
  set_identity_rows_kernel<CPLX><<<nbi, 256, 0, st>>>(Qt, m, k, ldt);
  for (int j = k - 1; j >= 0; --j) {
    const int nb = (int)ceil_div(k - j, Q_CPB);
    house_applyq_kernel<CPLX><<<nb, Q_THREADS, 0, st>>>(Qt, m, k, ldt, j, V, tau);
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
  extract_r_kernel<CPLX><<<nbr, 256, 0, st>>>(At, rdiag, nt, k, ldt, (T*)R, ldr, lq);
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
