// MPO-site application: the small-index contraction between the two big GEMMs of H_eff*C and of
// the environment update.
//
//   out[x, d, y1, f, y2] = sum_{p,q} W[p, d, q, f] * in[x, p, q, y]      y = y1*Y2 + y2
//
// W (the MPO site tensor, real, tiny and sparse) is given in CSR form over the output pair (d,f):
// entries (p*Q+q, value).  `in` and `out` are addressed with general element strides so that the
// same kernel serves the left/right environments and the MPS / MPDM (ancilla) index orders.
// HBM-bound: one read of `in`, one write of `out`; the (p,q) slab for a tile of y is staged in
// shared memory so that global loads and stores stay coalesced whichever index is contiguous.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

namespace rn {


template <bool CPLX>
__global__ void __launch_bounds__(256) wapply_kernel(WApplyParams p) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  extern __shared__ __align__(16) unsigned char w_smem_raw[];
  T* s_in = reinterpret_cast<T*>(w_smem_raw);
  const T* in = reinterpret_cast<const T*>(p.in) + (long)blockIdx.x * p.isx;
  T* out = reinterpret_cast<T*>(p.out) + (long)blockIdx.x * p.osx;
  const int y0 = blockIdx.y * p.YT;
  const int YT = p.YT, YTP = p.YT + 1, PQ = p.P * p.Q;
  const int nload = PQ * YT;
  for (int idx = threadIdx.x; idx < nload; idx += blockDim.x) {
    int pp, qq, y;
    if (p.order == 0) {
      y = idx % YT; const int pq = idx / YT; pp = pq / p.Q; qq = pq % p.Q;
    } else if (p.order == 1) {
      pp = idx % p.P; const int rest = idx / p.P; y = rest % YT; qq = rest / YT;
    } else {
      qq = idx % p.Q; const int rest = idx / p.Q; y = rest % YT; pp = rest / YT;
    }
    T v;
    if constexpr (CPLX) v = make_double2(0.0, 0.0); else v = 0.0;
    if (y0 + y < p.Y) v = in[(long)pp * p.isp + (long)qq * p.isq + (long)(y0 + y) * p.isy];
    s_in[(pp * p.Q + qq) * YTP + y] = v;
  }
  __syncthreads();
  const int nout = p.D * p.F * YT;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    const int y = o % YT, df = o / YT;
    const int yy = y0 + y;
    if (yy >= p.Y) continue;
    const int e0 = p.rowptr[df], e1 = p.rowptr[df + 1];
    T acc;
    if constexpr (CPLX) acc = make_double2(0.0, 0.0); else acc = 0.0;
    for (int e = e0; e < e1; ++e) {
      const double w = p.ent_val[e];
      const T v = s_in[p.ent_pq[e] * YTP + y];
      if constexpr (CPLX) { acc.x = fma(w, v.x, acc.x); acc.y = fma(w, v.y, acc.y); }
      else acc = fma(w, v, acc);
    }
    const int d = df / p.F, f = df % p.F;
    out[(long)d * p.osd + (long)f * p.osf + (long)(yy / p.Y2) * p.osy1 + (long)(yy % p.Y2) * p.osy2] = acc;
  }
}

// Wide MPO bonds (ab initio Hamiltonians: D*F and P*Q in the hundreds, ~50 entries per CSR row).  The
// kernel above gives one thread one output element, so every entry (value, column) is fetched from
// global memory once per element and the fetch latency (L2: the CSR arrays are larger than L1) sits on
// the critical path of every FMA: 92 ms per application at M = 1024, w = 326.  Here a thread owns a
// whole output row (d, f) of the y tile: an entry is fetched once and applied to the WW_YT y values the
// thread keeps in registers, the (p, q) slab is staged in shared memory as above: 73 ms.  (A
// warp-per-row variant -- lanes over y, entries handed round with shuffles, conflict-free 256-byte
// shared-memory reads, one block per SM -- measured slower, 107 ms.)  The bound is the shared-memory
// traffic of a sparse product with no reuse, nnz * Y * X * 8 bytes = 570 GB per application, 16 ms at
// the aggregate shared-memory bandwidth; what is left above that is latency at low occupancy.
constexpr int WW_YT = 16;

template <bool CPLX>
__global__ void __launch_bounds__(256) wapply_wide_kernel(WApplyParams p) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  extern __shared__ __align__(16) unsigned char w_smem_raw[];
  T* s_in = reinterpret_cast<T*>(w_smem_raw);
  const T* in = reinterpret_cast<const T*>(p.in) + (long)blockIdx.x * p.isx;
  T* out = reinterpret_cast<T*>(p.out) + (long)blockIdx.x * p.osx;
  constexpr int YT = WW_YT, YTP = WW_YT + 1;
  const int y0 = blockIdx.y * YT;
  const int PQ = p.P * p.Q;
  for (int idx = threadIdx.x; idx < PQ * YT; idx += blockDim.x) {
    int pp, qq, y;
    if (p.order == 0) {
      y = idx % YT; const int pq = idx / YT; pp = pq / p.Q; qq = pq % p.Q;
    } else if (p.order == 1) {
      pp = idx % p.P; const int rest = idx / p.P; y = rest % YT; qq = rest / YT;
    } else {
      qq = idx % p.Q; const int rest = idx / p.Q; y = rest % YT; pp = rest / YT;
    }
    T v;
    if constexpr (CPLX) v = make_double2(0.0, 0.0); else v = 0.0;
    if (y0 + y < p.Y) v = in[(long)pp * p.isp + (long)qq * p.isq + (long)(y0 + y) * p.isy];
    s_in[(pp * p.Q + qq) * YTP + y] = v;
  }
  __syncthreads();
  const int ndf = p.D * p.F;
  for (int df = threadIdx.x; df < ndf; df += blockDim.x) {
    const int e0 = p.rowptr[df], e1 = p.rowptr[df + 1];
    T acc[YT];
#pragma unroll
    for (int y = 0; y < YT; ++y) { if constexpr (CPLX) acc[y] = make_double2(0.0, 0.0); else acc[y] = 0.0; }
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
      const double w = p.ent_val[e];
      const T* row = s_in + p.ent_pq[e] * YTP;
#pragma unroll
      for (int y = 0; y < YT; ++y) {
        const T v = row[y];
        if constexpr (CPLX) { acc[y].x = fma(w, v.x, acc[y].x); acc[y].y = fma(w, v.y, acc[y].y); }
        else acc[y] = fma(w, v, acc[y]);
      }
    }
    const int d = df / p.F, f = df % p.F;
    T* orow = out + (long)d * p.osd + (long)f * p.osf;
#pragma unroll
    for (int y = 0; y < YT; ++y) {
      const int yy = y0 + y;
      if (yy < p.Y) orow[(long)(yy / p.Y2) * p.osy1 + (long)(yy % p.Y2) * p.osy2] = acc[y];
    }
  }
}

int launch_wapply(cudaStream_t st, int cplx, const WApplyParams& in_p) {
  WApplyParams p = in_p;
  if (p.X <= 0 || p.Y <= 0 || p.D <= 0 || p.F <= 0) return 0;
  {
    // wide bond: enough output rows to give every thread one, and the slab of a 16-wide y tile fits
    const int wide_yt = WW_YT;
    const long wide_smem = (long)p.P * p.Q * (WW_YT + 1) * (cplx ? 16 : 8);
    static int wide_on = -1;                // RN_WAPPLY_WIDE=0 keeps the element-per-thread kernel (diagnostics)
    if (wide_on < 0) { const char* e = getenv("RN_WAPPLY_WIDE"); wide_on = (e && e[0] == '0') ? 0 : 1; }
    if (wide_on && (long)p.D * p.F >= 128 && wide_smem <= 200 * 1024 && p.Y >= wide_yt) {
      p.YT = wide_yt;
      if (p.isy == 1) p.order = 0;
      else if (p.isp == 1) p.order = 1;
      else if (p.isq == 1) p.order = 2;
      else p.order = 0;
      static long wide_set[2] = {0, 0};
      if (wide_smem > 48 * 1024 && wide_smem > wide_set[cplx]) {
        if (cplx)
          RN_CHECK(cudaFuncSetAttribute(wapply_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        else
          RN_CHECK(cudaFuncSetAttribute(wapply_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        wide_set[cplx] = 200 * 1024;
      }
      dim3 grid((unsigned)p.X, (unsigned)ceil_div(p.Y, wide_yt));
      if (cplx) { RN_LAUNCH(wapply_wide_kernel<true>, grid, 256, (size_t)wide_smem, st, p); rn::g_launches++; }
      else { RN_LAUNCH(wapply_wide_kernel<false>, grid, 256, (size_t)wide_smem, st, p); rn::g_launches++; }
      RN_LAUNCH_CHECK();
      return 0;
    }
  }
  const int elt = cplx ? 16 : 8;
  const long budget = 160 * 1024;
  int yt = 128;
  while (yt > 1 && ((long)p.P * p.Q * (yt + 1) * elt > budget || yt / 2 >= p.Y)) yt >>= 1;
  if ((long)p.P * p.Q * (yt + 1) * elt > 200 * 1024) return (int)cudaErrorInvalidValue;
  p.YT = yt;
  if (p.isy == 1) p.order = 0;
  else if (p.isp == 1) p.order = 1;
  else if (p.isq == 1) p.order = 2;
  else p.order = 0;
  const int smem = p.P * p.Q * (yt + 1) * elt;
  static int max_set[2] = {0, 0};
  if (smem > 48 * 1024 && smem > max_set[cplx]) {
    if (cplx)
      RN_CHECK(cudaFuncSetAttribute(wapply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    else
      RN_CHECK(cudaFuncSetAttribute(wapply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    max_set[cplx] = 200 * 1024;
  }
  dim3 grid((unsigned)p.X, (unsigned)ceil_div(p.Y, yt));
  if (cplx) { RN_LAUNCH(wapply_kernel<true>, grid, 256, smem, st, p); rn::g_launches++; }
  else { RN_LAUNCH(wapply_kernel<false>, grid, 256, smem, st, p); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

}  // namespace rn

extern "C" int rn_wapply(void* stream, int cplx, const void* in, void* out, int X, int P, int Q,
                         int Y, long isx, long isp, long isq, long isy, int D, int F, int Y2,
                         long osx, long osd, long osf, long osy1, long osy2, const int* rowptr,
                         const int* ent_pq, const double* ent_val) {
  rn::WApplyParams p;
  p.in = in; p.out = out; p.X = X; p.P = P; p.Q = Q; p.Y = Y;
  p.isx = isx; p.isp = isp; p.isq = isq; p.isy = isy;
  p.D = D; p.F = F; p.Y2 = Y2;
  p.osx = osx; p.osd = osd; p.osf = osf; p.osy1 = osy1; p.osy2 = osy2;
  p.rowptr = rowptr; p.ent_pq = ent_pq; p.ent_val = ent_val; p.YT = 0; p.order = 0;
  return rn::launch_wapply((cudaStream_t)stream, cplx, p);
}
