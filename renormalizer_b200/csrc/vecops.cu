// Krylov / Davidson vector kernels: deterministic two-stage reductions and fused updates whose
// scalar coefficients stay on the device (no host round trip inside a Lanczos step).
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

namespace rn {

constexpr int V_THREADS = 256;

// partial[(i*NB + b)*2 + {0,1}] = sum over this block's slice of conj(V_i[k]) * x[k]
template <bool CPLX>
__global__ void __launch_bounds__(V_THREADS)
multi_dot_partial_kernel(const double* __restrict__ V, long ld, const double* __restrict__ x,
                         long n, double* __restrict__ partial) {
  pdl_wait();
  __shared__ double scratch[64];
  const double* v = V + (long)blockIdx.y * ld;
  double acc[2] = {0.0, 0.0};
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
    if constexpr (CPLX) {
      const double2 a = reinterpret_cast<const double2*>(v)[k];
      const double2 b = reinterpret_cast<const double2*>(x)[k];
      acc[0] += a.x * b.x + a.y * b.y;
      acc[1] += a.x * b.y - a.y * b.x;
    } else {
      acc[0] += v[k] * x[k];
    }
  }
  block_sum<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 0] = acc[0];
    partial[((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = acc[1];
  }
}

// out[2i + {0,1}] = sum_b partial[i][b]; post == 1: out[2i] = sqrt(re), out[2i+1] = 0
__global__ void __launch_bounds__(32)
reduce_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out, int post) {
  pdl_wait();
  double re = 0.0, im = 0.0;
  for (int b = threadIdx.x; b < nb; b += 32) {
    re += partial[((long)blockIdx.x * nb + b) * 2 + 0];
    im += partial[((long)blockIdx.x * nb + b) * 2 + 1];
  }
  re = warp_sum(re);
  im = warp_sum(im);
  if (threadIdx.x == 0) {
    if (post == 1) { re = sqrt(re > 0.0 ? re : 0.0); im = 0.0; }
    out[2 * blockIdx.x + 0] = re;
    out[2 * blockIdx.x + 1] = im;
  }
}

// w -= alpha * vj + beta * vjm1 on the double view; partial sums of w^2
__global__ void __launch_bounds__(V_THREADS)
lanczos_update_kernel(long nd, double* __restrict__ w, const double* __restrict__ vj,
                      const double* __restrict__ vjm1, const double* __restrict__ alpha_ptr,
                      const double* __restrict__ beta_ptr, double* __restrict__ partial) {
  pdl_wait();
  __shared__ double scratch[64];
  const double alpha = alpha_ptr[0];
  const double beta = (beta_ptr != nullptr && vjm1 != nullptr) ? beta_ptr[0] : 0.0;
  double acc[2] = {0.0, 0.0};
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += step) {
    double t = w[k] - alpha * vj[k];
    if (vjm1 != nullptr) t -= beta * vjm1[k];
    w[k] = t;
    acc[0] += t * t;
  }
  block_sum<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[(long)blockIdx.x * 2 + 0] = acc[0];
    partial[(long)blockIdx.x * 2 + 1] = 0.0;
  }
}

// out = x / s[0]
__global__ void __launch_bounds__(V_THREADS)
scale_inv_kernel(long nd, const double* __restrict__ x, const double* __restrict__ s,
                 double* __restrict__ out) {
  pdl_wait();
  const double inv = 1.0 / s[0];
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += step) out[k] = x[k] * inv;
}

// out[k] = sum_i coef[i] * V[i*ld + k]
template <bool CPLX>
__global__ void __launch_bounds__(V_THREADS)
lincomb_kernel(long n, int nvec, const double* __restrict__ V, long ld,
               const double* __restrict__ coef, double* __restrict__ out) {
  pdl_wait();
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
    if constexpr (CPLX) {
      double2 acc = make_double2(0.0, 0.0);
      for (int i = 0; i < nvec; ++i) {
        const double2 c = reinterpret_cast<const double2*>(coef)[i];
        const double2 v = reinterpret_cast<const double2*>(V + (long)i * ld)[k];
        acc.x += c.x * v.x - c.y * v.y;
        acc.y += c.x * v.y + c.y * v.x;
      }
      reinterpret_cast<double2*>(out)[k] = acc;
    } else {
      double acc = 0.0;
      for (int i = 0; i < nvec; ++i) acc += coef[i] * V[(long)i * ld + k];
      out[k] = acc;
    }
  }
}

// violations += #{k : !(|a_k - b_k| <= atol + rtol |b_k|)}   (numpy.allclose semantics, NaN counts)
template <bool CPLX>
__global__ void __launch_bounds__(V_THREADS)
allclose_kernel(long n, const double* __restrict__ a, const double* __restrict__ b, double rtol,
                double atol, int* __restrict__ violations) {
  pdl_wait();
  int bad = 0;
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
    double diff, mag;
    if constexpr (CPLX) {
      const double2 x = reinterpret_cast<const double2*>(a)[k];
      const double2 y = reinterpret_cast<const double2*>(b)[k];
      diff = hypot(x.x - y.x, x.y - y.y);
      mag = hypot(y.x, y.y);
    } else {
      diff = fabs(a[k] - b[k]);
      mag = fabs(b[k]);
    }
    if (!(diff <= atol + rtol * mag)) bad = 1;
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0 && bad) atomicAdd(violations, 1);
}

static inline int nblocks_for(long n) {
  long nb = ceil_div(n, (long)V_THREADS * 4);
  if (nb < 1) nb = 1;
  if (nb > RN_REDUCE_BLOCKS) nb = RN_REDUCE_BLOCKS;
  return (int)nb;
}

}  // namespace rn

using namespace rn;

extern "C" int rn_multi_dot(void* stream, int cplx, long n, int nvec, const double* V, long ld,
                            const double* x, double* ws, double* out) {
  if (nvec <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = nblocks_for(n);
  dim3 grid(nb, nvec);
  if (cplx) { RN_LAUNCH(multi_dot_partial_kernel<true>, grid, V_THREADS, 0, st, V, ld, x, n, ws); rn::g_launches++; }
  else { RN_LAUNCH(multi_dot_partial_kernel<false>, grid, V_THREADS, 0, st, V, ld, x, n, ws); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  { RN_LAUNCH(reduce_final_kernel, nvec, 32, 0, st, ws, nb, out, 0); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

extern "C" int rn_lanczos_update(void* stream, long nd, double* w, const double* vj,
                                 const double* vjm1, const double* alpha, const double* beta_prev,
                                 double* ws, double* beta_out) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = nblocks_for(nd);
  { RN_LAUNCH(lanczos_update_kernel, nb, V_THREADS, 0, st, nd, w, vj, vjm1, alpha, beta_prev, ws); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  { RN_LAUNCH(reduce_final_kernel, 1, 32, 0, st, ws, nb, beta_out, 1); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

extern "C" int rn_scale_inv(void* stream, long nd, const double* x, const double* s, double* out) {
  cudaStream_t st = (cudaStream_t)stream;
  int nb = nblocks_for(nd) * 4;
  { RN_LAUNCH(scale_inv_kernel, nb, V_THREADS, 0, st, nd, x, s, out); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

extern "C" int rn_lincomb(void* stream, int cplx, long n, int nvec, const double* V, long ld,
                          const double* coef, double* out) {
  cudaStream_t st = (cudaStream_t)stream;
  int nb = (int)ceil_div(n, V_THREADS);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  if (cplx) { RN_LAUNCH(lincomb_kernel<true>, nb, V_THREADS, 0, st, n, nvec, V, ld, coef, out); rn::g_launches++; }
  else { RN_LAUNCH(lincomb_kernel<false>, nb, V_THREADS, 0, st, n, nvec, V, ld, coef, out); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

extern "C" int rn_allclose(void* stream, int cplx, long n, const double* a, const double* b, double rtol,
                           double atol, int* violations) {
  cudaStream_t st = (cudaStream_t)stream;
  RN_CHECK(cudaMemsetAsync(violations, 0, sizeof(int), st));
  if (n <= 0) return 0;
  int nb = nblocks_for(n) * 2;
  if (cplx) { RN_LAUNCH(allclose_kernel<true>, nb, V_THREADS, 0, st, n, a, b, rtol, atol, violations); rn::g_launches++; }
  else { RN_LAUNCH(allclose_kernel<false>, nb, V_THREADS, 0, st, n, a, b, rtol, atol, violations); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}
