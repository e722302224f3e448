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
#include "qr_common.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace rn {

constexpr int Q_THREADS = 1024;  // passes over a column are L2-latency bound: few rows per thread
constexpr int QT_THREADS = 256;   // register-heavy T builder
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


// ---- single-launch variants ------------------------------------------------------------------
// Elimination of all k reflectors in ONE cooperative launch: column c is owned by block
// c % gridDim.x for the whole factorisation, a grid barrier separates the steps (column j+1 must
// be final before every block derives reflector j+1 from it).
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_factor_coop_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, int k, long ldt,
                         typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                         double* __restrict__ rdiag) {
  pdl_wait();
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
  pdl_wait();
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


// ---- dataflow variant: no grid barrier -------------------------------------------------------
// Column c becomes "final" once reflectors 0..c-1 have been applied to it; its owner then
// publishes (|a_c[c+1:]|^2, a_c[c]) and a ready flag.  Every block waits only for the flag of
// the reflector it needs next, so the critical path is owner(j+1): wait ready[j] -> dot ->
// update -> publish ready[j+1], with no chip-wide synchronisation.  Launched cooperatively only
// for the co-residency guarantee (a waiting block must never keep a producer from being scheduled).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool CPLX>
__device__ __forceinline__ typename Cx<CPLX>::T ldcg_elt(const typename Cx<CPLX>::T* p) {
  if constexpr (CPLX) return __ldcg(reinterpret_cast<const double2*>(p));
  else return __ldcg(reinterpret_cast<const double*>(p));
}

template <bool CPLX>
__device__ __forceinline__ void reflector_from_info(const double* info, typename Cx<CPLX>::T& tau,
                                                    typename Cx<CPLX>::T& scale, double& beta) {
  using C = Cx<CPLX>;
  const double ss = __ldcg(info + 0), ar = __ldcg(info + 1), ai = __ldcg(info + 2);
  if (ss == 0.0 && ai == 0.0) { tau = C::zero(); scale = C::zero(); beta = ar; return; }
  const double nrm = sqrt(ar * ar + ai * ai + ss);
  beta = ar >= 0.0 ? -nrm : nrm;
  tau = C::make((beta - ar) / beta, -ai / beta);
  const double dr = ar - beta, di = ai, den = dr * dr + di * di;
  scale = C::make(dr / den, -di / den);
}

// the owner of a freshly final column c (< k) publishes it and stores its reflector
template <bool CPLX>
__device__ __forceinline__ void publish_column(typename Cx<CPLX>::T* At, int m, long ldt, int c, double ssq,
                                               double are, double aim, double* colinfo, int* ready,
                                               typename Cx<CPLX>::T* V, typename Cx<CPLX>::T* tau_out,
                                               double* rdiag) {
  using C = Cx<CPLX>;
  using T = typename C::T;
  if (threadIdx.x == 0) {
    colinfo[4 * c + 0] = ssq; colinfo[4 * c + 1] = are; colinfo[4 * c + 2] = aim;
    __threadfence();
    st_release_gpu(ready + c, 1);
  }
  // off the critical path: V_c, tau_c, R_cc
  T tau, scale;
  double beta;
  if (ssq == 0.0 && aim == 0.0) { tau = C::zero(); scale = C::zero(); beta = are; }
  else {
    const double nrm = sqrt(are * are + aim * aim + ssq);
    beta = are >= 0.0 ? -nrm : nrm;
    tau = C::make((beta - are) / beta, -aim / beta);
    const double dr = are - beta, di = aim, den = dr * dr + di * di;
    scale = C::make(dr / den, -di / den);
  }
  const T* col = At + (long)c * ldt;
  T* vc = V + (long)c * ldt;
  for (int r = threadIdx.x; r < m; r += blockDim.x)
    vc[r] = r < c ? C::zero() : (r == c ? C::one() : C::mul(col[r], scale));
  if (threadIdx.x == 0) { tau_out[c] = tau; rdiag[c] = beta; }
}

template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_factor_flow_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, int k, long ldt,
                         typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                         double* __restrict__ rdiag, double* __restrict__ colinfo, int* __restrict__ ready) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_CPB * 32];
  const int nb = gridDim.x, bid = blockIdx.x;
  if (bid == 0) {
    // column 0 is final from the start
    const T* c0 = At;
    double ss[1] = {0.0};
    for (int r = 1 + threadIdx.x; r < m; r += blockDim.x) ss[0] += C::abs2(c0[r]);
    block_sum<1>(ss, scratch);
    publish_column<CPLX>(At, m, ldt, 0, ss[0], C::re(c0[0]), C::im(c0[0]), colinfo, ready, V, tau_out, rdiag);
    __syncthreads();
  }
  for (int j = 0; j < k; ++j) {
    int c = j + 1 + ((bid - (j + 1)) % nb + nb) % nb;     // first owned column > j
    if (c >= n) break;                                     // nothing left for this block
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(ready + j) == 0) { }
    }
    __syncthreads();
    const T* colj = At + (long)j * ldt;
    T tau, scale;
    double beta;
    reflector_from_info<CPLX>(colinfo + 4 * j, tau, scale, beta);
    const T ctau = C::conj(tau);
    // the next reflector source first, alone, so that its flag goes out as early as possible
    if (c == j + 1) {
      T* col = At + (long)c * ldt;
      double acc[2] = {0.0, 0.0};
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : C::mul(ldcg_elt<CPLX>(colj + r), scale);
        const T d = C::cmul(v, col[r]);
        acc[0] += C::re(d); acc[1] += C::im(d);
      }
      block_sum<2>(acc, scratch);
      const T w = C::mul(ctau, C::make(acc[0], acc[1]));
      double nfo[3] = {0.0, 0.0, 0.0};
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : C::mul(ldcg_elt<CPLX>(colj + r), scale);
        const T x = C::sub(col[r], C::mul(v, w));
        col[r] = x;
        if (r > c) nfo[0] += C::abs2(x);
        else if (r == c) { nfo[1] = C::re(x); nfo[2] = C::im(x); }
      }
      if (c < k) {
        block_sum<3>(nfo, scratch);
        publish_column<CPLX>(At, m, ldt, c, nfo[0], nfo[1], nfo[2], colinfo, ready, V, tau_out, rdiag);
      }
      __syncthreads();
      c += nb;
    }
    while (c < n) {
      int cols[Q_CPB];
#pragma unroll
      for (int i = 0; i < Q_CPB; ++i) { cols[i] = c < n ? c : -1; c += nb; }
      double acc[2 * Q_CPB];
#pragma unroll
      for (int i = 0; i < 2 * Q_CPB; ++i) acc[i] = 0.0;
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : C::mul(ldcg_elt<CPLX>(colj + r), scale);
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
        const T v = r == j ? C::one() : C::mul(ldcg_elt<CPLX>(colj + r), scale);
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
  }
}


// Same flag-chained elimination with the block's own columns (and the current reflector) held in
// SHARED memory: the per-step critical path is then one L2 read of the pivot column, two
// shared-memory passes and one global write of the freshly final column.
template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_factor_flow_smem_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, int k, long ldt,
                              typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                              double* __restrict__ rdiag, double* __restrict__ colinfo,
                              int* __restrict__ ready, int cpb) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
  __shared__ double scratch[4 * 32];
  T* own = reinterpret_cast<T*>(qr_smem_raw);          // [cpb][m]
  T* vbuf = own + (long)cpb * m;                       // [m]
  const int nb = gridDim.x, bid = blockIdx.x;
  int nown = 0;
  for (int c = bid; c < n; c += nb, ++nown)
    for (int r = threadIdx.x; r < m; r += blockDim.x) own[(long)nown * m + r] = At[(long)c * ldt + r];
  __syncthreads();
  if (bid == 0) {
    double ss[1] = {0.0};
    for (int r = 1 + threadIdx.x; r < m; r += blockDim.x) ss[0] += C::abs2(own[r]);
    block_sum<1>(ss, scratch);
    publish_column<CPLX>(At, m, ldt, 0, ss[0], C::re(own[0]), C::im(own[0]), colinfo, ready, V, tau_out, rdiag);
    __syncthreads();
  }
  for (int j = 0; j < k; ++j) {
    // local index of the first owned column > j
    int l0 = (j + 1 - bid + nb - 1) / nb;
    if (j + 1 <= bid) l0 = 0;
    if (l0 >= nown) break;
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(ready + j) == 0) { }
    }
    __syncthreads();
    const T* colj = At + (long)j * ldt;
    T tau, scale;
    double beta;
    reflector_from_info<CPLX>(colinfo + 4 * j, tau, scale, beta);
    const T ctau = C::conj(tau);
    for (int r = j + threadIdx.x; r < m; r += blockDim.x)
      vbuf[r] = r == j ? C::one() : C::mul(ldcg_elt<CPLX>(colj + r), scale);
    __syncthreads();
    for (int l = l0; l < nown; ++l) {
      const int c = bid + l * nb;
      T* col = own + (long)l * m;
      double acc[2] = {0.0, 0.0};
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T d = C::cmul(vbuf[r], col[r]);
        acc[0] += C::re(d); acc[1] += C::im(d);
      }
      block_sum<2>(acc, scratch);
      const T w = C::mul(ctau, C::make(acc[0], acc[1]));
      const bool becomes_final = (c == j + 1) && (c < k);
      double nfo[3] = {0.0, 0.0, 0.0};
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T x = C::sub(col[r], C::mul(vbuf[r], w));
        col[r] = x;
        if (becomes_final) {
          if (r > c) nfo[0] += C::abs2(x);
          else if (r == c) { nfo[1] = C::re(x); nfo[2] = C::im(x); }
        }
      }
      if (becomes_final) {
        __syncthreads();
        T* gcol = At + (long)c * ldt;
        for (int r = threadIdx.x; r < m; r += blockDim.x) gcol[r] = col[r];
        block_sum<3>(nfo, scratch);
        publish_column<CPLX>(At, m, ldt, c, nfo[0], nfo[1], nfo[2], colinfo, ready, V, tau_out, rdiag);
      }
      __syncthreads();
    }
  }
  // columns that never become reflector sources (c >= k, wide matrices) and, for safety, every
  // owned column: the shared copy is the truth
  for (int l = 0; l < nown; ++l) {
    const int c = bid + l * nb;
    if (c >= k)
      for (int r = threadIdx.x; r < m; r += blockDim.x) At[(long)c * ldt + r] = own[(long)l * m + r];
  }
}


// ---- panel variant ----------------------------------------------------------------------------
// Q_PB adjacent columns form a panel owned by one block and held in shared memory.  The owner
// applies the reflectors of every earlier panel (Q_PB per flag hand-off, read from V), factors
// its own panel locally, writes R / V / tau back and raises the panel's flag: n / Q_PB chained
// hand-offs instead of n.
constexpr int Q_PB = 4;

template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_factor_panel_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, int n, int k, long ldt,
                          typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                          double* __restrict__ rdiag, int* __restrict__ ready) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  extern __shared__ __align__(16) unsigned char qr_smem_raw[];
  __shared__ double scratch[2 * Q_PB * 32];
  __shared__ T s_tau[Q_PB];
  T* own = reinterpret_cast<T*>(qr_smem_raw);          // [Q_PB][m]
  const int p = blockIdx.x;
  const int c0 = p * Q_PB;
  if (c0 >= n) return;
  const int ncol = (n - c0) < Q_PB ? (n - c0) : Q_PB;
  for (int a = 0; a < ncol; ++a)
    for (int r = threadIdx.x; r < m; r += blockDim.x) own[(long)a * m + r] = At[(long)(c0 + a) * ldt + r];
  __syncthreads();
  // reflectors of the earlier panels
  for (int q = 0; q < p; ++q) {
    const int j0 = q * Q_PB;
    if (j0 >= k) break;
    const int nrefl = (k - j0) < Q_PB ? (k - j0) : Q_PB;
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(ready + q) == 0) { }
    }
    __syncthreads();
    if (threadIdx.x < nrefl) s_tau[threadIdx.x] = ldcg_elt<CPLX>(tau_out + j0 + threadIdx.x);
    for (int a = 0; a < nrefl; ++a) {
      const int j = j0 + a;
      const T* vj = V + (long)j * ldt;
      double acc[2 * Q_PB];
#pragma unroll
      for (int i = 0; i < 2 * Q_PB; ++i) acc[i] = 0.0;
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = ldcg_elt<CPLX>(vj + r);
#pragma unroll
        for (int i = 0; i < Q_PB; ++i)
          if (i < ncol) {
            const T d = C::cmul(v, own[(long)i * m + r]);
            acc[2 * i] += C::re(d);
            acc[2 * i + 1] += C::im(d);
          }
      }
      block_sum<2 * Q_PB>(acc, scratch);       // also publishes s_tau
      const T ctau = C::conj(s_tau[a]);
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = ldcg_elt<CPLX>(vj + r);
#pragma unroll
        for (int i = 0; i < Q_PB; ++i)
          if (i < ncol) {
            const T w = C::mul(ctau, C::make(acc[2 * i], acc[2 * i + 1]));
            own[(long)i * m + r] = C::sub(own[(long)i * m + r], C::mul(v, w));
          }
      }
      __syncthreads();
    }
  }
  // factor the own panel
  for (int a = 0; a < ncol; ++a) {
    const int j = c0 + a;
    if (j >= k) break;
    T* col = own + (long)a * m;
    double nfo[3] = {0.0, 0.0, 0.0};
    for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
      const T x = col[r];
      if (r > j) nfo[0] += C::abs2(x);
      else { nfo[1] = C::re(x); nfo[2] = C::im(x); }
    }
    block_sum<3>(nfo, scratch);
    T tau, scale;
    double beta;
    {
      const double ss = nfo[0], ar = nfo[1], ai = nfo[2];
      if (ss == 0.0 && ai == 0.0) { tau = C::zero(); scale = C::zero(); beta = ar; }
      else {
        const double nrm = sqrt(ar * ar + ai * ai + ss);
        beta = ar >= 0.0 ? -nrm : nrm;
        tau = C::make((beta - ar) / beta, -ai / beta);
        const double dr = ar - beta, di = ai, den = dr * dr + di * di;
        scale = C::make(dr / den, -di / den);
      }
    }
    // v_j -> V (global) and, scaled in place, below the diagonal of the shared column
    T* vj = V + (long)j * ldt;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
      T v;
      if (r < j) v = C::zero();
      else if (r == j) v = C::one();
      else { v = C::mul(col[r], scale); col[r] = v; }
      vj[r] = v;
    }
    if (threadIdx.x == 0) { tau_out[j] = tau; rdiag[j] = beta; }
    __syncthreads();
    if (a + 1 < ncol) {
      double acc[2 * Q_PB];
#pragma unroll
      for (int i = 0; i < 2 * Q_PB; ++i) acc[i] = 0.0;
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : col[r];
#pragma unroll
        for (int i = 0; i < Q_PB; ++i)
          if (i > a && i < ncol) {
            const T d = C::cmul(v, own[(long)i * m + r]);
            acc[2 * i] += C::re(d);
            acc[2 * i + 1] += C::im(d);
          }
      }
      block_sum<2 * Q_PB>(acc, scratch);
      const T ctau = C::conj(tau);
      for (int r = j + threadIdx.x; r < m; r += blockDim.x) {
        const T v = r == j ? C::one() : col[r];
#pragma unroll
        for (int i = 0; i < Q_PB; ++i)
          if (i > a && i < ncol) {
            const T w = C::mul(ctau, C::make(acc[2 * i], acc[2 * i + 1]));
            own[(long)i * m + r] = C::sub(own[(long)i * m + r], C::mul(v, w));
          }
      }
      __syncthreads();
    }
  }
  // R entries (rows <= column index) back to global; the part below the diagonal is not read again
  for (int a = 0; a < ncol; ++a) {
    const int c = c0 + a;
    const int rmax = c < m ? c + 1 : m;
    for (int r = threadIdx.x; r < rmax; r += blockDim.x) At[(long)c * ldt + r] = own[(long)a * m + r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release_gpu(ready + p, 1);
  }
}

// Q = H_0 ... H_{k-1} I, one WARP per column of Q (no block-level synchronisation): column c
// needs reflectors c..0 only.  Reflectors are shared by all warps and stay L1/L2 resident.
template <bool CPLX>
__global__ void __launch_bounds__(256)
house_formq_warp_kernel(typename Cx<CPLX>::T* __restrict__ Qt, int m, int k, long ldt,
                        const typename Cx<CPLX>::T* __restrict__ V,
                        const typename Cx<CPLX>::T* __restrict__ tau_in) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= k) return;
  T* q = Qt + (long)c * ldt;
  for (int r = lane; r < m; r += 32) q[r] = (r == c) ? C::one() : C::zero();
  __syncwarp();
  for (int j = c; j >= 0; --j) {
    const T* vj = V + (long)j * ldt;
    double dre = 0.0, dim = 0.0;
    for (int r = j + lane; r < m; r += 32) {
      const T d = C::cmul(vj[r], q[r]);
      dre += C::re(d); dim += C::im(d);
    }
    dre = warp_sum(dre);
    dim = warp_sum(dim);
    const T w = C::mul(tau_in[j], C::make(dre, dim));
    for (int r = j + lane; r < m; r += 32) q[r] = C::sub(q[r], C::mul(vj[r], w));
    __syncwarp();
  }
}


// ---- blocked Q formation (compact WY, groups of Q_GB reflectors) -----------------------------
// H_{j0} ... H_{j0+nb-1} = I - V T V^H (LAPACK larft, forward / columnwise).  One small kernel
// builds every group's T from the Gram matrix of its reflectors; the Q kernel then applies a
// whole group per pass (nb dot products, one tiny triangular product, one update), i.e. Q_GB
// times fewer latency-bound phases than reflector-by-reflector application.
constexpr int Q_GB = 8;
constexpr int Q_QCOLS = 1;

template <bool CPLX>
__global__ void __launch_bounds__(QT_THREADS)
house_build_t_kernel(const typename Cx<CPLX>::T* __restrict__ V, const typename Cx<CPLX>::T* __restrict__ tau_in,
                     int m, int k, long ldt, typename Cx<CPLX>::T* __restrict__ Tall) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_GB * Q_GB * 32 / 2];   // 28 pairs * 2 <= 64 values
  __shared__ T G[Q_GB][Q_GB];
  __shared__ T Ts[Q_GB][Q_GB];
  const int j0 = blockIdx.x * Q_GB;
  const int nb = (k - j0) < Q_GB ? (k - j0) : Q_GB;
  // Gram of the group: G[a][b] = v_a^H v_b, a < b  (28 pairs -> 56 doubles)
  double acc[2 * 28];
#pragma unroll
  for (int i = 0; i < 2 * 28; ++i) acc[i] = 0.0;
  for (int r = j0 + threadIdx.x; r < m; r += blockDim.x) {
    T v[Q_GB];
#pragma unroll
    for (int a = 0; a < Q_GB; ++a) v[a] = a < nb ? V[(long)(j0 + a) * ldt + r] : C::zero();
    int p = 0;
#pragma unroll
    for (int a = 0; a < Q_GB; ++a)
#pragma unroll
      for (int b = a + 1; b < Q_GB; ++b, ++p) {
        const T d = C::cmul(v[a], v[b]);
        acc[2 * p] += C::re(d);
        acc[2 * p + 1] += C::im(d);
      }
  }
  block_sum<2 * 28>(acc, scratch);
  if (threadIdx.x == 0) {
    int p = 0;
    for (int a = 0; a < Q_GB; ++a)
      for (int b = a + 1; b < Q_GB; ++b, ++p) G[a][b] = C::make(acc[2 * p], acc[2 * p + 1]);
    for (int a = 0; a < Q_GB; ++a)
      for (int b = 0; b < Q_GB; ++b) Ts[a][b] = C::zero();
    for (int b = 0; b < nb; ++b) {
      const T tb = tau_in[j0 + b];
      Ts[b][b] = tb;
      // T[0:b, b] = -tau_b * T[0:b, 0:b] * G[0:b, b]
      for (int a = 0; a < b; ++a) {
        T sacc = C::zero();
        for (int c = a; c < b; ++c) {
          const T t = C::mul(Ts[a][c], G[c][b]);
          sacc = C::make(C::re(sacc) + C::re(t), C::im(sacc) + C::im(t));
        }
        const T t2 = C::mul(tb, sacc);
        Ts[a][b] = C::make(-C::re(t2), -C::im(t2));
      }
    }
    for (int a = 0; a < Q_GB; ++a)
      for (int b = 0; b < Q_GB; ++b) Tall[((long)blockIdx.x * Q_GB + a) * Q_GB + b] = Ts[a][b];
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(Q_THREADS)
house_formq_blocked_kernel(typename Cx<CPLX>::T* __restrict__ Qt, int m, int k, long ldt,
                           const typename Cx<CPLX>::T* __restrict__ V,
                           const typename Cx<CPLX>::T* __restrict__ Tall) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  __shared__ double scratch[2 * Q_GB * Q_QCOLS * 32];
  __shared__ T Ts[Q_GB][Q_GB];
  const int c0 = blockIdx.x * Q_QCOLS;
  if (c0 >= k) return;
  const int ncol = (k - c0) < Q_QCOLS ? (k - c0) : Q_QCOLS;
  for (int i = 0; i < ncol; ++i)
    for (int r = threadIdx.x; r < m; r += blockDim.x)
      Qt[(long)(c0 + i) * ldt + r] = (r == c0 + i) ? C::one() : C::zero();
  __syncthreads();
  for (int g = (c0 + ncol - 1) / Q_GB; g >= 0; --g) {
    const int j0 = g * Q_GB;
    const int nb = (k - j0) < Q_GB ? (k - j0) : Q_GB;
    if (threadIdx.x < Q_GB * Q_GB)
      Ts[threadIdx.x / Q_GB][threadIdx.x % Q_GB] = Tall[(long)g * Q_GB * Q_GB + threadIdx.x];
    // w[a][i] = v_a^H q_i
    double acc[2 * Q_GB * Q_QCOLS];
#pragma unroll
    for (int i = 0; i < 2 * Q_GB * Q_QCOLS; ++i) acc[i] = 0.0;
    for (int r = j0 + threadIdx.x; r < m; r += blockDim.x) {
      T q[Q_QCOLS];
#pragma unroll
      for (int i = 0; i < Q_QCOLS; ++i) q[i] = i < ncol ? Qt[(long)(c0 + i) * ldt + r] : C::zero();
#pragma unroll
      for (int a = 0; a < Q_GB; ++a) {
        const T v = a < nb ? V[(long)(j0 + a) * ldt + r] : C::zero();
#pragma unroll
        for (int i = 0; i < Q_QCOLS; ++i) {
          const T d = C::cmul(v, q[i]);
          acc[2 * (a * Q_QCOLS + i)] += C::re(d);
          acc[2 * (a * Q_QCOLS + i) + 1] += C::im(d);
        }
      }
    }
    block_sum<2 * Q_GB * Q_QCOLS>(acc, scratch);     // also orders the Ts stores
    // z = T w
    T z[Q_GB][Q_QCOLS];
#pragma unroll
    for (int a = 0; a < Q_GB; ++a)
#pragma unroll
      for (int i = 0; i < Q_QCOLS; ++i) {
        T sacc = C::zero();
#pragma unroll
        for (int b = 0; b < Q_GB; ++b) {
          if (b >= a) {
            const T t = C::mul(Ts[a][b], C::make(acc[2 * (b * Q_QCOLS + i)], acc[2 * (b * Q_QCOLS + i) + 1]));
            sacc = C::make(C::re(sacc) + C::re(t), C::im(sacc) + C::im(t));
          }
        }
        z[a][i] = sacc;
      }
    for (int r = j0 + threadIdx.x; r < m; r += blockDim.x) {
      T upd[Q_QCOLS];
#pragma unroll
      for (int i = 0; i < Q_QCOLS; ++i) upd[i] = C::zero();
#pragma unroll
      for (int a = 0; a < Q_GB; ++a) {
        const T v = a < nb ? V[(long)(j0 + a) * ldt + r] : C::zero();
#pragma unroll
        for (int i = 0; i < Q_QCOLS; ++i) {
          const T t = C::mul(v, z[a][i]);
          upd[i] = C::make(C::re(upd[i]) + C::re(t), C::im(upd[i]) + C::im(t));
        }
      }
#pragma unroll
      for (int i = 0; i < Q_QCOLS; ++i)
        if (i < ncol) {
          T* p = Qt + (long)(c0 + i) * ldt + r;
          *p = C::sub(*p, upd[i]);
        }
    }
    __syncthreads();
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
static int g_qr_warp_formq = 0;
static int g_qr_use_flow = 1;   // flag-chained elimination (slower on B200 than the grid barrier)
static int g_qr_coop_blocks[2] = {-1, -1};   // co-resident block budget per dtype, -1 = unknown

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
  int& budget = g_qr_coop_blocks[CPLX ? 1 : 0];
  if (budget < 0) {
    if (const char* e = getenv("RN_QR_FLOW")) g_qr_use_flow = atoi(e);
    if (const char* e = getenv("RN_QR_WARP_FORMQ")) g_qr_warp_formq = atoi(e);
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    RN_CHECK(cudaGetDevice(&dev));
    RN_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    RN_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, house_factor_flow_kernel<CPLX>,
                                                           Q_THREADS, 0));
    budget = coop ? sms * (per_sm > 1 ? 1 : per_sm) : 0;
  }
  if (budget > 0) {
    int nb = n < budget ? n : budget;
    if (nb < 1) nb = 1;
    double* colinfo = nullptr;
    int* ready = nullptr;
    RN_CHECK(cudaMallocAsync((void**)&colinfo, sizeof(double) * 4 * (size_t)n, st));
    RN_CHECK(cudaMallocAsync((void**)&ready, sizeof(int) * (size_t)n, st));
    RN_CHECK(cudaMemsetAsync(ready, 0, sizeof(int) * (size_t)n, st));
    void* args[] = {(void*)&At, (void*)&m, (void*)&n, (void*)&k, (void*)&ldt, (void*)&V, (void*)&tau,
                    (void*)&rdiag, (void*)&colinfo, (void*)&ready};
    const int cpb = (int)ceil_div(n, nb);
    const size_t smem_need = (size_t)(cpb + 1) * m * sizeof(typename Cx<CPLX>::T);
    const int npanels = (int)ceil_div(n, Q_PB);
    const size_t smem_panel = (size_t)Q_PB * m * sizeof(typename Cx<CPLX>::T);
    if (g_qr_use_flow == 4 && npanels <= budget && smem_panel <= 200 * 1024) {
      static bool attr_done[2] = {false, false};
      if (!attr_done[CPLX ? 1 : 0]) {
        RN_CHECK(cudaFuncSetAttribute(house_factor_panel_kernel<CPLX>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done[CPLX ? 1 : 0] = true;
      }
      void* args3[] = {(void*)&At, (void*)&m, (void*)&n, (void*)&k, (void*)&ldt, (void*)&V, (void*)&tau,
                       (void*)&rdiag, (void*)&ready};
      { RN_CHECK(cudaLaunchCooperativeKernel((void*)house_factor_panel_kernel<CPLX>, dim3(npanels),
                                             dim3(Q_THREADS), args3, smem_panel, st)); rn::g_launches++; }
    } else if (g_qr_use_flow == 1 && smem_need <= 200 * 1024) {
      static bool attr_done[2] = {false, false};
      if (!attr_done[CPLX ? 1 : 0]) {
        RN_CHECK(cudaFuncSetAttribute(house_factor_flow_smem_kernel<CPLX>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done[CPLX ? 1 : 0] = true;
      }
      int cpb_arg = cpb;
      void* args2[] = {(void*)&At, (void*)&m, (void*)&n, (void*)&k, (void*)&ldt, (void*)&V, (void*)&tau,
                       (void*)&rdiag, (void*)&colinfo, (void*)&ready, (void*)&cpb_arg};
      { RN_CHECK(cudaLaunchCooperativeKernel((void*)house_factor_flow_smem_kernel<CPLX>, dim3(nb), dim3(Q_THREADS),
                                           args2, smem_need, st)); rn::g_launches++; }
    } else if (g_qr_use_flow)
      { RN_CHECK(cudaLaunchCooperativeKernel((void*)house_factor_flow_kernel<CPLX>, dim3(nb), dim3(Q_THREADS),
                                           args, 0, st)); rn::g_launches++; }
    else
      { RN_CHECK(cudaLaunchCooperativeKernel((void*)house_factor_coop_kernel<CPLX>, dim3(nb), dim3(Q_THREADS),
                                           args, 0, st)); rn::g_launches++; }
    if (g_qr_warp_formq == 1)
      { RN_LAUNCH(house_formq_warp_kernel<CPLX>, (unsigned)ceil_div(k, 8), 256, 0, st, Qt, m, k, ldt, V, tau); rn::g_launches++; }
    else if (g_qr_warp_formq == 2)
      { RN_LAUNCH(house_formq_kernel<CPLX>, (unsigned)ceil_div(k, Q_CPB), Q_THREADS, 0, st, Qt, m, k, ldt, V, tau); rn::g_launches++; }
    else {
      typename Cx<CPLX>::T* Tall = nullptr;
      const int ngroups = (int)ceil_div(k, Q_GB);
      RN_CHECK(cudaMallocAsync((void**)&Tall, sizeof(typename Cx<CPLX>::T) * (size_t)ngroups * Q_GB * Q_GB, st));
      { RN_LAUNCH(house_build_t_kernel<CPLX>, ngroups, QT_THREADS, 0, st, V, tau, m, k, ldt, Tall); rn::g_launches++; }
      { RN_LAUNCH(house_formq_blocked_kernel<CPLX>, (unsigned)ceil_div(k, Q_QCOLS), Q_THREADS, 0, st, Qt, m, k, ldt, V, Tall); rn::g_launches++; }
      RN_CHECK(cudaFreeAsync(Tall, st));
    }
    RN_LAUNCH_CHECK();
    RN_CHECK(cudaFreeAsync(colinfo, st));
    RN_CHECK(cudaFreeAsync(ready, st));
    return 0;
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
