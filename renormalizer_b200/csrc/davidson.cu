// Davidson eigensolver for the lowest eigenpair of H_eff, the whole iteration driven from C++ with the
// trial vectors, their images and every vector operation on the device (reference:
// renormalizer/lib/davidson/davidson.py:73-455, PySCF's davidson1 with the reference's switches, for
// nroots == 1; the DMRG caller is renormalizer/mps/gs.py:486-576).
//
// Per iteration the stream sees
//   H_eff apply (hop.cu)  ->  <ax_i, x_s> partials -> reduce -> [host: Rayleigh matrix, 12x12 eigh]
//   -> residual + preconditioner fused over both stacks (coefficients passed by value), |r|^2, |t|^2
//      partials -> reduce -> [host: convergence test]
//   -> <x_i, t> partials -> reduce -> projection + |t'|^2 partials -> x_{s+1} = t' / |t'|
// The host sees two small scalar blocks per iteration through pinned, device-mapped memory (the last
// kernel of each phase writes them followed by an epoch word; no copy node, no stream synchronise):
// the Python version of round 1 read six device scalars per iteration with a blocking copy each.
// The subspace schedule, the convergence rule (|de| < tol and |r| < sqrt(tol)), the restart rule
// and the linear-dependency thresholds are the reference's, so the number of H_eff applications is
// the reference's (asserted by the tests).
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

#include <math.h>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

namespace rn {

constexpr int D_THREADS = 256;
constexpr int D_MAXS = 24;            // largest subspace the by-value coefficient block holds

struct DavCoef {                      // complex coefficients of the Ritz vector, by value
  double re[D_MAXS], im[D_MAXS];
};

static inline int d_nblocks(long n) {
  long nb = ceil_div(n, (long)D_THREADS * 4);
  if (nb < 1) nb = 1;
  if (nb > RN_REDUCE_BLOCKS) nb = RN_REDUCE_BLOCKS;
  return (int)nb;
}

// partial[(i*NB + b)*2 + {0,1}] = this block's share of <V_i, x> * scale (complex: conj(V_i) x)
template <bool CPLX>
__global__ void __launch_bounds__(D_THREADS)
dav_dots_kernel(const double* __restrict__ V, long ld, const double* __restrict__ x, long n,
                double* __restrict__ partial) {
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

// out[2i + {0,1}] = sum_b partial[i][b] (fixed order); when `host` is given the values are also
// published to mapped host memory at host[offset + 2i ..] and, by the last block, the epoch word.
__global__ void __launch_bounds__(32)
dav_reduce_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out,
                  volatile double* __restrict__ host, int* __restrict__ ticket, volatile int* __restrict__ host_epoch,
                  int epoch) {
  pdl_wait();
  double re = 0.0, im = 0.0;
  for (int b = threadIdx.x; b < nb; b += 32) {
    re += partial[((long)blockIdx.x * nb + b) * 2 + 0];
    im += partial[((long)blockIdx.x * nb + b) * 2 + 1];
  }
  re = warp_sum(re);
  im = warp_sum(im);
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x + 0] = re;
    out[2 * blockIdx.x + 1] = im;
    if (host != nullptr) {
      host[2 * blockIdx.x + 0] = re;
      host[2 * blockIdx.x + 1] = im;
      __threadfence_system();
      if (host_epoch != nullptr) {                     // quiet publish otherwise: read at the next epoch
        const int t = atomicAdd(ticket, 1);
        if (t == (int)gridDim.x - 1) {
          *ticket = 0;
          __threadfence_system();
          *host_epoch = epoch;
        }
      }
    }
  }
}

// y = (y + add) * scale   (sum over the members of a stacked Hamiltonian, `inverse` factor)
__global__ void __launch_bounds__(D_THREADS)
dav_axpy_scale_kernel(long nd, double* __restrict__ y, const double* __restrict__ add, double scale) {
  pdl_wait();
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += step)
    y[k] = (y[k] + (add != nullptr ? add[k] : 0.0)) * scale;
}

// Ritz vector, residual and preconditioned correction in one pass over both stacks:
//   x0 = sum_i c_i XS_i,  ax0 = sum_i c_i AX_i,  r = mask (ax0 - e x0),  t = mask r / (hdiag - e + 1e-4)
// (davidson.py:383-405 and gs.py:512-514; the quantum-number mask is applied here instead of inside
// H_eff: the trial vectors never leave the allowed subspace, so the Rayleigh matrix is unchanged).
// partial[b] = (|r|^2, |t|^2) of this block.  When x0_out is given the Ritz vector is stored as well.
template <bool CPLX>
__global__ void __launch_bounds__(D_THREADS)
dav_residual_kernel(long n, int nvec, const double* __restrict__ XS, const double* __restrict__ AX, long ld,
                    DavCoef c, double e, const unsigned char* __restrict__ mask, const double* __restrict__ hdiag,
                    double* __restrict__ t_out, double* __restrict__ x0_out, double* __restrict__ partial) {
  pdl_wait();
  __shared__ double scratch[64];
  double acc[2] = {0.0, 0.0};
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
    const bool ok = mask == nullptr || mask[k] != 0;
    if constexpr (CPLX) {
      double2 x0 = make_double2(0.0, 0.0), ax0 = make_double2(0.0, 0.0);
      for (int i = 0; i < nvec; ++i) {
        const double2 xv = reinterpret_cast<const double2*>(XS + (long)i * ld)[k];
        const double2 av = reinterpret_cast<const double2*>(AX + (long)i * ld)[k];
        x0.x += c.re[i] * xv.x - c.im[i] * xv.y; x0.y += c.re[i] * xv.y + c.im[i] * xv.x;
        ax0.x += c.re[i] * av.x - c.im[i] * av.y; ax0.y += c.re[i] * av.y + c.im[i] * av.x;
      }
      if (x0_out != nullptr) reinterpret_cast<double2*>(x0_out)[k] = x0;
      double2 r = make_double2(ax0.x - e * x0.x, ax0.y - e * x0.y);
      if (!ok) r = make_double2(0.0, 0.0);
      const double inv = ok ? 1.0 / (hdiag[k] - e + 1e-4) : 0.0;
      const double2 t = make_double2(r.x * inv, r.y * inv);
      if (t_out != nullptr) reinterpret_cast<double2*>(t_out)[k] = t;
      acc[0] += r.x * r.x + r.y * r.y;
      acc[1] += t.x * t.x + t.y * t.y;
    } else {
      double x0 = 0.0, ax0 = 0.0;
      for (int i = 0; i < nvec; ++i) {
        x0 += c.re[i] * XS[(long)i * ld + k];
        ax0 += c.re[i] * AX[(long)i * ld + k];
      }
      if (x0_out != nullptr) x0_out[k] = x0;
      const double r = ok ? ax0 - e * x0 : 0.0;
      const double t = ok ? r / (hdiag[k] - e + 1e-4) : 0.0;
      if (t_out != nullptr) t_out[k] = t;
      acc[0] += r * r;
      acc[1] += t * t;
    }
  }
  block_sum<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[(long)blockIdx.x * 2 + 0] = acc[0];
    partial[(long)blockIdx.x * 2 + 1] = acc[1];
  }
}

// t' = t * tscale - sum_i (dots_i * tscale) XS_i   (davidson.py:407-411 with t normalised first);
// partial[b] = (|t'|^2, 0).  dots = <XS_i, t> of the un-normalised t (device array of pairs).
template <bool CPLX>
__global__ void __launch_bounds__(D_THREADS)
dav_project_kernel(long n, int nvec, const double* __restrict__ XS, long ld, const double* __restrict__ dots,
                   double tscale, double* __restrict__ t, double* __restrict__ partial) {
  pdl_wait();
  __shared__ double scratch[64];
  __shared__ double cr[D_MAXS], ci[D_MAXS];
  if (threadIdx.x < nvec) { cr[threadIdx.x] = dots[2 * threadIdx.x] * tscale; ci[threadIdx.x] = dots[2 * threadIdx.x + 1] * tscale; }
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
    if constexpr (CPLX) {
      double2 v = reinterpret_cast<double2*>(t)[k];
      v.x *= tscale; v.y *= tscale;
      for (int i = 0; i < nvec; ++i) {
        const double2 xv = reinterpret_cast<const double2*>(XS + (long)i * ld)[k];
        v.x -= cr[i] * xv.x - ci[i] * xv.y;
        v.y -= cr[i] * xv.y + ci[i] * xv.x;
      }
      reinterpret_cast<double2*>(t)[k] = v;
      acc[0] += v.x * v.x + v.y * v.y;
    } else {
      double v = t[k] * tscale;
      for (int i = 0; i < nvec; ++i) v -= cr[i] * XS[(long)i * ld + k];
      t[k] = v;
      acc[0] += v * v;
    }
  }
  block_sum<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[(long)blockIdx.x * 2 + 0] = acc[0];
    partial[(long)blockIdx.x * 2 + 1] = 0.0;
  }
}

// out = x / sqrt(nrm2[0])  (0 when the norm vanishes)
__global__ void __launch_bounds__(D_THREADS)
dav_normalise_kernel(long nd, const double* __restrict__ x, const double* __restrict__ nrm2, double* __restrict__ out) {
  pdl_wait();
  const double s = nrm2[0];
  const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
  const long step = (long)gridDim.x * blockDim.x;
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += step) out[k] = x[k] * inv;
}

// ---- host side: cyclic Jacobi for the (at most D_MAXS x D_MAXS) Hermitian Rayleigh matrix -------
typedef std::complex<double> cd;
static void herm_eig_lowest(const std::vector<cd>& hin, int n, double* e_out, cd* v_out) {
  std::vector<cd> a(hin), v((size_t)n * n, cd(0.0));
  for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < n; ++p) {
      diag += std::norm(a[(size_t)p * n + p]);
      for (int q = p + 1; q < n; ++q) off += std::norm(a[(size_t)p * n + q]);
    }
    if (off <= 1e-34 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const cd apq = a[(size_t)p * n + q];
        const double g = std::abs(apq);
        if (g == 0.0) continue;
        const double app = a[(size_t)p * n + p].real(), aqq = a[(size_t)q * n + q].real();
        const double zeta = (aqq - app) / (2.0 * g);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        const cd ph = apq / g;                         // a_pq = g ph
        // columns: a[:,p] <- c a[:,p] - s conj(ph) a[:,q] ; a[:,q] <- s ph a[:,p] + c a[:,q]
        for (int k = 0; k < n; ++k) {
          const cd akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * std::conj(ph) * akq;
          a[(size_t)k * n + q] = s * ph * akp + c * akq;
          const cd vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * vkp - s * std::conj(ph) * vkq;
          v[(size_t)k * n + q] = s * ph * vkp + c * vkq;
        }
        // rows (J^H from the left)
        for (int k = 0; k < n; ++k) {
          const cd apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * ph * aqk;
          a[(size_t)q * n + k] = s * std::conj(ph) * apk + c * aqk;
        }
        a[(size_t)p * n + q] = 0.0; a[(size_t)q * n + p] = 0.0;
        a[(size_t)p * n + p] = a[(size_t)p * n + p].real();
        a[(size_t)q * n + q] = a[(size_t)q * n + q].real();
      }
  }
  int best = 0;
  for (int i = 1; i < n; ++i)
    if (a[(size_t)i * n + i].real() < a[(size_t)best * n + best].real()) best = i;
  *e_out = a[(size_t)best * n + best].real();
  for (int k = 0; k < n; ++k) v_out[k] = v[(size_t)k * n + best];
}

namespace {
struct DavFlags {
  volatile double* host = nullptr;     // [0 .. 2*D_MAXS): scalar block, then the epoch word (as int)
  double* dev = nullptr;
  int epoch = 0;
};
std::mutex g_df_mu;
std::map<cudaStream_t, DavFlags> g_df;
DavFlags& dav_flags(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_df_mu);
  return g_df[st];
}
}  // namespace

}  // namespace rn

using namespace rn;

extern "C" int rn_davidson(rn_hop_plan** plans, int nplans, void* stream, int cplx, long n, const void* x0_in,
                           const unsigned char* mask, const double* hdiag, double inverse, double tol,
                           int max_cycle, int max_space, double lindep, void* c_out, double* e_out,
                           int* nhop_out, int* converged_out) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0 || nplans < 1 || plans == nullptr) return (int)cudaErrorInvalidValue;
  if (max_space + 2 > D_MAXS) return (int)cudaErrorInvalidValue;
  const int es = cplx ? 2 : 1;
  const long nd = n * es;
  const int cap = max_space + 2;
  const int nb = d_nblocks(n);
  int nbw = (int)ceil_div(nd, (long)D_THREADS * 2);
  if (nbw > 148 * 8) nbw = 148 * 8;
  if (nbw < 1) nbw = 1;
  const double toloose = sqrt(tol);

  double *XS = nullptr, *AX = nullptr, *tvec = nullptr, *tmp = nullptr, *small = nullptr;
  auto cleanup = [&]() {
    if (XS) cudaFreeAsync(XS, st);
    if (AX) cudaFreeAsync(AX, st);
    if (tvec) cudaFreeAsync(tvec, st);
    if (tmp) cudaFreeAsync(tmp, st);
    if (small) cudaFreeAsync(small, st);
  };
#define DAV_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { cleanup(); return (int)_e; } } while (0)
#define DAV_TRY(x) do { int _r = (x); if (_r) { cleanup(); return _r; } } while (0)
  DAV_CUDA(cudaMallocAsync((void**)&XS, sizeof(double) * (size_t)nd * cap, st));
  DAV_CUDA(cudaMallocAsync((void**)&AX, sizeof(double) * (size_t)nd * cap, st));
  DAV_CUDA(cudaMallocAsync((void**)&tvec, sizeof(double) * (size_t)nd, st));
  if (nplans > 1) DAV_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * (size_t)nd, st));
  const size_t partial_doubles = (size_t)2 * RN_REDUCE_BLOCKS * (cap + 1);
  DAV_CUDA(cudaMallocAsync((void**)&small, sizeof(double) * (partial_doubles + 2 * (cap + 2) + 4) + 64, st));
  double* partial = small;
  double* red = partial + partial_doubles;             // reduced scalars (pairs)
  int* ticket = reinterpret_cast<int*>(red + 2 * (cap + 2));
  DAV_CUDA(cudaMemsetAsync(ticket, 0, sizeof(int), st));

  DavFlags& df = dav_flags(st);
  if (!df.host) {
    double* hp = nullptr;
    DAV_CUDA(cudaHostAlloc((void**)&hp, sizeof(double) * (2 * D_MAXS + 2), cudaHostAllocMapped));
    for (int i = 0; i < 2 * D_MAXS + 2; ++i) hp[i] = 0.0;
    DAV_CUDA(cudaHostGetDevicePointer((void**)&df.dev, hp, 0));
    df.host = hp;
  }
  volatile double* hbuf = df.host;
  volatile int* h_epoch = reinterpret_cast<volatile int*>(df.host + 2 * D_MAXS);
  double* dbuf = df.dev;
  int* d_epoch = reinterpret_cast<int*>(df.dev + 2 * D_MAXS);
  int& epoch = df.epoch;

  auto wait_epoch = [&]() -> int {
    long spins = 0;
    while (*h_epoch != epoch) {
      if ((++spins & 0xfffff) == 0) {
        cudaError_t q = cudaStreamQuery(st);
        if (q != cudaSuccess && q != cudaErrorNotReady) return (int)q;
        if (q == cudaSuccess && *h_epoch != epoch) return (int)cudaErrorUnknown;
      }
    }
    return 0;
  };
  // reduce `count` pairs of partial sums and publish them to the host block
  auto reduce_publish = [&](int count, int nblocks) {
    ++epoch;
    RN_LAUNCH(dav_reduce_kernel, count, 32, 0, st, (const double*)partial, nblocks, red, (volatile double*)dbuf, ticket,
              (volatile int*)d_epoch, epoch);
    rn::g_launches++;
  };
  auto reduce_device = [&](int count, int nblocks) {
    RN_LAUNCH(dav_reduce_kernel, count, 32, 0, st, (const double*)partial, nblocks, red, (volatile double*)nullptr,
              ticket, (volatile int*)nullptr, 0);
    rn::g_launches++;
  };
  // one pair, reduced on the device and written to the deferred slot of the host block (no epoch):
  // the host reads it at its next synchronisation point
  constexpr int DEFER = 2 * (D_MAXS - 1);
  auto reduce_deferred = [&](int nblocks) {
    RN_LAUNCH(dav_reduce_kernel, 1, 32, 0, st, (const double*)partial, nblocks, red, (volatile double*)(dbuf + DEFER),
              ticket, (volatile int*)nullptr, 0);
    rn::g_launches++;
  };
  auto dots = [&](const double* V, int nvec, const double* x) {
    if (cplx) RN_LAUNCH(dav_dots_kernel<true>, dim3(nb, nvec), D_THREADS, 0, st, V, nd, x, n, partial);
    else RN_LAUNCH(dav_dots_kernel<false>, dim3(nb, nvec), D_THREADS, 0, st, V, nd, x, n, partial);
    rn::g_launches++;
  };
  auto apply_h = [&](const double* x, double* y) -> int {
    int err = rn_hop_apply(plans[0], st, x, y);
    for (int p = 1; p < nplans && !err; ++p) {
      err = rn_hop_apply(plans[p], st, x, tmp);
      if (err) break;
      RN_LAUNCH(dav_axpy_scale_kernel, nbw, D_THREADS, 0, st, nd, y, (const double*)tmp, p == nplans - 1 ? inverse : 1.0);
      rn::g_launches++;
    }
    if (!err && nplans == 1 && inverse != 1.0) {
      RN_LAUNCH(dav_axpy_scale_kernel, nbw, D_THREADS, 0, st, nd, y, (const double*)nullptr, inverse);
      rn::g_launches++;
    }
    return err;
  };

  std::vector<cd> heff((size_t)cap * cap, cd(0.0)), sub;
  std::vector<cd> v(cap, cd(0.0));
  DavCoef coef, coef_prev;
  double e = 0.0, elast = 0.0, e_prev = 0.0;
  int space = 0, space_prev = 0, nhop = 0, converged = 0;
  bool fresh = true, have_v = false, deferred = false;
  const double* start = (const double*)x0_in;
  for (int icyc = 0; icyc < max_cycle; ++icyc) {
    if (fresh) {
      // davidson.py:335-342 (_qr of the start vector / of the current Ritz vector on a restart):
      // normalise, reject a vanishing vector
      dots(start, 1, start);
      reduce_publish(1, nb);
      RN_LAUNCH(dav_normalise_kernel, nbw, D_THREADS, 0, st, nd, start, (const double*)red, XS);
      rn::g_launches++;
      DAV_CUDA(cudaGetLastError());
      DAV_TRY(wait_epoch());
      if (deferred && !(hbuf[DEFER] > lindep)) break;  // the correction of the last cycle was dependent
      deferred = false;
      if (!(hbuf[0] > lindep)) {
        if (!have_v) { cleanup(); return (int)cudaErrorInvalidValue; }   // "initial guess is empty or zero"
        break;
      }
      space = 0;
    }
    double* xs = XS + (long)space * nd;
    double* ax = AX + (long)space * nd;
    DAV_TRY(apply_h(xs, ax));
    // new row of the Rayleigh matrix: d_i = <ax_i, x_j>, heff[j,i] = conj(d_i)  (davidson.py:56-70)
    dots(AX, space + 1, xs);
    reduce_publish(space + 1, nb);
    DAV_CUDA(cudaGetLastError());
    DAV_TRY(wait_epoch());
    if (deferred && !(hbuf[DEFER] > lindep)) {
      // the projected correction was linearly dependent on the subspace: the reference stops before
      // applying H_eff to it (davidson.py:413-420); this application is discarded and not counted
      coef = coef_prev; e = e_prev; space = space_prev;
      break;
    }
    deferred = false;
    ++nhop;
    ++space;
    const int j = space - 1;
    for (int i = 0; i < space; ++i) {
      const cd d(hbuf[2 * i], cplx ? hbuf[2 * i + 1] : 0.0);
      heff[(size_t)j * cap + i] = std::conj(d);
      heff[(size_t)i * cap + j] = d;
    }
    heff[(size_t)j * cap + j] = heff[(size_t)j * cap + j].real();
    sub.assign((size_t)space * space, cd(0.0));
    for (int a = 0; a < space; ++a)
      for (int b = 0; b < space; ++b) sub[(size_t)a * space + b] = heff[(size_t)a * cap + b];
    elast = e;
    herm_eig_lowest(sub, space, &e, v.data());
    have_v = true;
    for (int i = 0; i < D_MAXS; ++i) { coef.re[i] = i < space ? v[i].real() : 0.0; coef.im[i] = i < space ? v[i].imag() : 0.0; }
    const double de = e - elast;
    // residual, preconditioned correction and their norms
    if (cplx) RN_LAUNCH(dav_residual_kernel<true>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, tvec, (double*)nullptr, partial);
    else RN_LAUNCH(dav_residual_kernel<false>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, tvec, (double*)nullptr, partial);
    rn::g_launches++;
    reduce_publish(1, nb);
    DAV_CUDA(cudaGetLastError());
    DAV_TRY(wait_epoch());
    const double r2 = hbuf[0], t2 = hbuf[1];
    const double rnorm = sqrt(r2 > 0.0 ? r2 : 0.0);
    if (fabs(de) < tol && rnorm < toloose) { converged = 1; break; }
    if (!(r2 > lindep)) break;                         // no correction vector left (davidson.py:389-396)
    if (icyc == max_cycle - 1) break;
    // normalised correction, projected on the complement of the subspace (davidson.py:398-411); its
    // norm test (davidson.py:413-420) is read by the host at the next synchronisation point
    const double tscale = t2 > 0.0 ? 1.0 / sqrt(t2) : 0.0;
    dots(XS, space, tvec);
    reduce_device(space, nb);
    if (cplx) RN_LAUNCH(dav_project_kernel<true>, nb, D_THREADS, 0, st, n, space, (const double*)XS, nd, (const double*)red, tscale, tvec, partial);
    else RN_LAUNCH(dav_project_kernel<false>, nb, D_THREADS, 0, st, n, space, (const double*)XS, nd, (const double*)red, tscale, tvec, partial);
    rn::g_launches++;
    reduce_deferred(nb);
    deferred = true;
    coef_prev = coef; e_prev = e; space_prev = space;
    fresh = space + 1 > max_space;                     // davidson.py:422 with nroots == 1
    if (!fresh) {
      RN_LAUNCH(dav_normalise_kernel, nbw, D_THREADS, 0, st, nd, (const double*)tvec, (const double*)red, XS + (long)space * nd);
      rn::g_launches++;
    } else {
      // restart from the current Ritz vector (the correction itself is dropped, davidson.py:335-342)
      if (cplx) RN_LAUNCH(dav_residual_kernel<true>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, (double*)nullptr, tvec, partial);
      else RN_LAUNCH(dav_residual_kernel<false>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, (double*)nullptr, tvec, partial);
      rn::g_launches++;
      start = tvec;
    }
    DAV_CUDA(cudaGetLastError());
  }
  // eigenvector: the Ritz vector of the last Rayleigh-Ritz step
  if (!have_v) { cleanup(); return (int)cudaErrorUnknown; }
  if (cplx) RN_LAUNCH(dav_residual_kernel<true>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, (double*)nullptr, (double*)c_out, partial);
  else RN_LAUNCH(dav_residual_kernel<false>, nb, D_THREADS, 0, st, n, space, (const double*)XS, (const double*)AX, nd, coef, e, mask, hdiag, (double*)nullptr, (double*)c_out, partial);
  rn::g_launches++;
  DAV_CUDA(cudaGetLastError());
  if (e_out) *e_out = e;
  if (nhop_out) *nhop_out = nhop;
  if (converged_out) *converged_out = converged;
  cleanup();
#undef DAV_CUDA
#undef DAV_TRY
  return 0;
}
