// Blocked Householder QR for the bond matrices of a sweep site (tall m x n, n up to a few
// hundred): LAPACK geqrf numerics (zlarfg reflectors, compact-WY block updates) laid out for
// Blackwell's thread-block clusters.
//
//   for every panel of QP_B columns:
//     1. house_panel_cluster_kernel  -- ONE cluster of QP_CS CTAs holds the panel in (distributed)
//        shared memory, rows split across the CTAs.  A reflector needs one pass of dot products
//        (the column norm, the products with the columns to its right for the update and with
//        the reflectors to its left for the T factor), ONE all-reduce of QP_B complex numbers
//        through DSMEM + one cluster barrier, and one update pass.  Nothing touches L2/HBM
//        between the initial load and the final store of the panel.
//     2. wy_dots_kernel / wy_update_kernel -- the compact-WY update of the trailing columns,
//        A <- (I - V T^H V^H) A, as two small tiled FP64 kernels over (row chunk, column group).
//   Q is then accumulated backwards, Q <- (I - V T V^H) Q, with the same two kernels.
//
// Storage is "column-as-row" (At[c*ldt + r] = A[r][c]) as in qr.cu, so every column is a
// contiguous stream.  V keeps the reflectors with their explicit unit diagonal; entries above
// the panel's first row are never read.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"
#include "qr_common.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace rn {

constexpr int QP_B = 32;          // panel width (reflectors per block reflector)
constexpr int QP_CS = 8;          // CTAs per cluster (portable maximum)
constexpr int QP_THREADS = 512;
constexpr int QP_RCH = 128;       // rows per chunk in the WY kernels
constexpr int QP_CGW = 32;        // columns per group in the WY kernels

template <bool CPLX>
__device__ __forceinline__ typename Cx<CPLX>::T cx_add(typename Cx<CPLX>::T a, typename Cx<CPLX>::T b) {
  return Cx<CPLX>::make(Cx<CPLX>::re(a) + Cx<CPLX>::re(b), Cx<CPLX>::im(a) + Cx<CPLX>::im(b));
}

// ---- 1. panel factorisation in one cluster -----------------------------------------------------
// Panel = columns [j0, j0+bw) of At, rows [j0, m); CTA `rank` owns local rows
// [j0 + rank*rloc, j0 + (rank+1)*rloc).  On exit: At holds R (on and above the diagonal, beta on
// it), V the reflectors (rows >= j0), tau/rdiag the scalars and Tout the bw x bw (stride QP_B)
// upper-triangular T of  H_j0 ... H_{j0+bw-1} = I - V T V^H.
template <bool CPLX>
__global__ void __launch_bounds__(QP_THREADS, 1)
house_panel_cluster_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, long ldt, int j0, int bw, int rloc,
                           typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                           double* __restrict__ rdiag, typename Cx<CPLX>::T* __restrict__ Tout) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char qp_smem_raw[];
  T* pan = reinterpret_cast<T*>(qp_smem_raw);                       // [bw][rloc]
  __shared__ __align__(16) unsigned char xbuf_raw[2 * QP_CS * QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char pbuf_raw[2 * QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char part_raw[QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char wv_raw[QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char tvec_raw[QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char ts_raw[QP_B * QP_B * sizeof(double2)];
  __shared__ __align__(16) unsigned char sc_raw[2 * sizeof(double2)];
  __shared__ double s_beta;
  T* xbuf = reinterpret_cast<T*>(xbuf_raw);       // [parity][src rank][column]
  T* pbuf = reinterpret_cast<T*>(pbuf_raw);       // [parity][column]: the pivot row
  T* part = reinterpret_cast<T*>(part_raw);
  T* wv = reinterpret_cast<T*>(wv_raw);
  T* tvec = reinterpret_cast<T*>(tvec_raw);
  T* Ts = reinterpret_cast<T*>(ts_raw);           // [i][l], stride QP_B
  T* s_sc = reinterpret_cast<T*>(sc_raw);         // [0] = tau, [1] = scale

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = j0 + rank * rloc;
  int nloc = m - r0;
  if (nloc > rloc) nloc = rloc;
  if (nloc < 0) nloc = 0;
  for (int c = warp; c < bw; c += QP_THREADS / 32) {
    const T* src = At + (long)(j0 + c) * ldt + r0;
    for (int i = lane; i < nloc; i += 32) pan[(long)c * rloc + i] = src[i];
  }
  for (int i = tid; i < QP_B * QP_B; i += QP_THREADS) Ts[i] = C::zero();
  __syncthreads();

  for (int jj = 0; jj < bw; ++jj) {
    const int par = jj & 1;
    const int owner = jj / rloc, pl = jj - owner * rloc;     // CTA and local row of the pivot
    const int vstart = rank == owner ? pl : (rank > owner ? 0 : nloc);
    const int lo = rank == owner ? pl + 1 : vstart;          // first local row strictly below it
    const T* pj = pan + (long)jj * rloc;
    // -- dot products of column jj (rows below the pivot) with every column of the panel
    for (int c = warp; c < bw; c += QP_THREADS / 32) {
      const T* pc = pan + (long)c * rloc;
      double ar = 0.0, ai = 0.0;
      for (int i = lo + lane; i < nloc; i += 32) {
        const T d = C::cmul(pj[i], pc[i]);
        ar += C::re(d); ai += C::im(d);
      }
      ar = warp_sum(ar); ai = warp_sum(ai);
      if (lane == 0) part[c] = C::make(ar, ai);
    }
    __syncthreads();
    // -- all-reduce through distributed shared memory
    if (tid < bw) {
      const T mine = part[tid];
      for (int rk = 0; rk < QP_CS; ++rk) {
        T* remote = cluster.map_shared_rank(xbuf, rk);
        remote[(par * QP_CS + rank) * QP_B + tid] = mine;
      }
      if (rank == owner) {
        const T pv = pan[(long)tid * rloc + pl];
        for (int rk = 0; rk < QP_CS; ++rk) {
          T* remote = cluster.map_shared_rank(pbuf, rk);
          remote[par * QP_B + tid] = pv;
        }
      }
    }
    cluster.sync();
    // -- reflector scalars (redundantly in every CTA / thread of warp 0) and the update vector
    if (warp == 0) {
      double ss = 0.0;
      for (int rk = 0; rk < QP_CS; ++rk) ss += C::re(xbuf[(par * QP_CS + rk) * QP_B + jj]);
      const T alpha = pbuf[par * QP_B + jj];
      const double ar = C::re(alpha), ai = C::im(alpha);
      T tau, scale;
      double beta;
      if (ss == 0.0 && ai == 0.0) { tau = C::zero(); scale = C::zero(); beta = ar; }
      else {
        const double nrm = sqrt(ar * ar + ai * ai + ss);
        beta = ar >= 0.0 ? -nrm : nrm;
        tau = C::make((beta - ar) / beta, -ai / beta);
        const double dr = ar - beta, di = ai, den = dr * dr + di * di;
        scale = C::make(dr / den, -di / den);
      }
      if (lane < bw) {
        T tot = C::zero();
        for (int rk = 0; rk < QP_CS; ++rk) tot = cx_add<CPLX>(tot, xbuf[(par * QP_CS + rk) * QP_B + lane]);
        const T pv = pbuf[par * QP_B + lane];
        if (lane > jj) {
          // a_c -= conj(tau) (v^H a_c) v,  v^H a_c = a_c[pivot] + conj(scale) sum conj(x) a_c
          wv[lane] = C::mul(C::conj(tau), cx_add<CPLX>(pv, C::mul(C::conj(scale), tot)));
        } else if (lane < jj) {
          // v_c^H v_jj = conj(v_c[pivot]) + scale * conj(sum conj(x) v_c)
          tvec[lane] = cx_add<CPLX>(C::conj(pv), C::mul(scale, C::conj(tot)));
        }
      }
      if (lane == 0) { s_sc[0] = tau; s_sc[1] = scale; s_beta = beta; }
      __syncwarp();
      // T[0:jj, jj] = -tau T[0:jj, 0:jj] (V^H v_jj),  T[jj, jj] = tau     (LAPACK larft)
      if (lane < jj) {
        T acc = C::zero();
        for (int l = lane; l < jj; ++l) acc = cx_add<CPLX>(acc, C::mul(Ts[lane * QP_B + l], tvec[l]));
        const T t2 = C::mul(tau, acc);
        Ts[lane * QP_B + jj] = C::make(-C::re(t2), -C::im(t2));
      } else if (lane == jj) {
        Ts[jj * QP_B + jj] = tau;
      }
    }
    __syncthreads();
    // -- update of the columns to the right (thread <-> row; two column groups)
    {
      const T scale = s_sc[1];
      const int half = tid >> 8, it = tid & 255;
      for (int i = vstart + it; i < nloc; i += 256) {
        const T v = (rank == owner && i == pl) ? C::one() : C::mul(pj[i], scale);
        for (int c = jj + 1 + half; c < bw; c += 2) {
          T* p = pan + (long)c * rloc + i;
          *p = C::sub(*p, C::mul(wv[c], v));
        }
      }
    }
    __syncthreads();
    // -- column jj becomes the reflector (below the pivot) and R's diagonal entry
    {
      const T scale = s_sc[1];
      T* pjw = pan + (long)jj * rloc;
      for (int i = lo + tid; i < nloc; i += QP_THREADS) pjw[i] = C::mul(pjw[i], scale);
      if (rank == owner && tid == 0) pjw[pl] = C::make(s_beta, 0.0);
      if (rank == 0 && tid == 0) { tau_out[j0 + jj] = s_sc[0]; rdiag[j0 + jj] = s_beta; }
    }
    __syncthreads();
  }
  // -- store: R part / reflectors
  for (int c = warp; c < bw; c += QP_THREADS / 32) {
    T* dstA = At + (long)(j0 + c) * ldt + r0;
    T* dstV = V + (long)(j0 + c) * ldt + r0;
    const int piv = j0 + c;
    for (int i = lane; i < nloc; i += 32) {
      const T val = pan[(long)c * rloc + i];
      const int r = r0 + i;
      dstA[i] = val;
      dstV[i] = r < piv ? C::zero() : (r == piv ? C::one() : val);
    }
  }
  if (rank == 0)
    for (int i = tid; i < QP_B * QP_B; i += QP_THREADS) Tout[i] = Ts[i];
  cluster.sync();      // no CTA may exit while a sibling can still write into its shared memory
}


// ---- DSMEM exchange primitives (st.async + mbarrier: no cluster-wide barrier per reflector) ------
__device__ __forceinline__ uint32_t qp_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void qp_st_async(uint32_t raddr, double2 v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "d"(v.x), "d"(v.y), "r"(rbar) : "memory");
}
__device__ __forceinline__ void qp_st_async(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
               ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void qp_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void qp_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qp_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "QPWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra QPDONE_%=;\n"
      "bra QPWAIT_%=;\n"
      "QPDONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- 1b. register-resident panel factorisation ---------------------------------------------------
// Same algorithm with the panel slice of each CTA held in REGISTERS: warp w owns columns
// w + 16 q (q < NCOL), lane l owns local rows l + 32 it (it < NIT).  Per reflector: the pivot
// column is broadcast through a small shared buffer, every warp forms the dot products of its
// own columns with it, pushes them to all CTAs of the cluster (DSMEM), and after ONE cluster
// barrier recomputes the reflector scalars redundantly (no block-level hand-off) and updates its
// own columns in registers.  Reflectors are kept unscaled (x instead of v = scale * x) until the
// final store, so a finished column is never written again.  One __syncthreads per reflector.
template <bool CPLX, int CS, int NIT, int NCOL>
__global__ void __launch_bounds__(512, 1)
house_panel_reg_kernel(typename Cx<CPLX>::T* __restrict__ At, int m, long ldt, int j0, int bw, int rloc,
                       typename Cx<CPLX>::T* __restrict__ V, typename Cx<CPLX>::T* __restrict__ tau_out,
                       double* __restrict__ rdiag, typename Cx<CPLX>::T* __restrict__ Tout) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  constexpr int BW = 16 * NCOL;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char qp_smem_raw[];
  constexpr int E = sizeof(double2);
  T* xcol = reinterpret_cast<T*>(qp_smem_raw);                       // pivot column broadcast
  T* xbuf = reinterpret_cast<T*>(qp_smem_raw + NIT * 32 * E);        // [parity][src rank][column]
  T* pbuf = reinterpret_cast<T*>(qp_smem_raw + (NIT * 32 + 2 * CS * BW) * E);            // [parity][column]: pivot row
  T* Ts = reinterpret_cast<T*>(qp_smem_raw + (NIT * 32 + 2 * CS * BW + 2 * BW) * E);     // [i][l], stride BW (rank 0)
  T* Gs = Ts + BW * BW;                           // [l][j] = v_l^H v_j, l < j (rank 0 only)
  T* Ws = Gs + BW * BW;
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = j0 + rank * rloc;
  int nloc = m - r0;
  if (nloc > rloc) nloc = rloc;
  if (nloc < 0) nloc = 0;

  T a[NCOL][NIT];
#pragma unroll
  for (int q = 0; q < NCOL; ++q) {
    const int c = warp + 16 * q;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = lane + 32 * it;
      a[q][it] = (c < bw && i < nloc) ? At[(long)(j0 + c) * ldt + r0 + i] : C::zero();
    }
  }
  for (int i = tid; i < BW * BW; i += 512) { Ts[i] = C::zero(); Gs[i] = C::zero(); }
  if (warp == 0) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) xcol[lane + 32 * it] = a[0][it];
  }
  if (tid == 0) {
    qp_mbar_init(smem_u32(&mbar[0]), 1);
    qp_mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();                                 // barriers initialised cluster-wide before any st.async

  T my_scale = C::zero(), my_tau = C::zero();     // lane c keeps the scalars of column c
  double my_beta = 0.0;
  const uint32_t xbuf_a = smem_u32(xbuf), pbuf_a = smem_u32(pbuf);
  int owner = 0, pl = 0;                          // CTA and local row of the pivot
  for (int jj = 0; jj < bw; ++jj, ++pl) {
    const int par = jj & 1;
    if (pl == rloc) { pl = 0; ++owner; }
    const uint32_t bar = smem_u32(&mbar[par]);
    if (tid == 0) qp_mbar_expect_tx(bar, (uint32_t)((CS + 1) * bw * sizeof(T)));
    T x[NIT];
    bool below[NIT], pivot_here[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = lane + 32 * it;
      x[it] = xcol[i];
      below[it] = i < nloc && (rank > owner || (rank == owner && i > pl));
      pivot_here[it] = rank == owner && i == pl;
    }
    // -- dot products of column jj (rows below the pivot) with this warp's columns
    T dots[NCOL], pvs[NCOL];
#pragma unroll
    for (int q = 0; q < NCOL; ++q) {
      double ar = 0.0, ai = 0.0, pr = 0.0, pi = 0.0;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        if (below[it]) {
          if constexpr (CPLX) {
            ar = fma(x[it].x, a[q][it].x, ar); ar = fma(x[it].y, a[q][it].y, ar);
            ai = fma(x[it].x, a[q][it].y, ai); ai = fma(-x[it].y, a[q][it].x, ai);
          } else {
            ar = fma(x[it], a[q][it], ar);
          }
        }
        if (pivot_here[it]) { pr = C::re(a[q][it]); pi = C::im(a[q][it]); }
      }
      ar = warp_sum(ar);
      if constexpr (CPLX) ai = warp_sum(ai);
      dots[q] = C::make(ar, ai);
      if (rank == owner) {                         // warp-uniform: broadcast the pivot-row entry
        pr = warp_sum(pr);
        if constexpr (CPLX) pi = warp_sum(pi);
      }
      pvs[q] = C::make(pr, pi);
    }
    // -- all-reduce through distributed shared memory: lane rk feeds CTA rk with st.async, the
    //    bytes complete the receiver's mbarrier (no cluster-wide barrier, no release fence)
    if (lane < CS) {
      const uint32_t rx = qp_mapa(xbuf_a, lane), rp = qp_mapa(pbuf_a, lane), rbar = qp_mapa(bar, lane);
#pragma unroll
      for (int q = 0; q < NCOL; ++q) {
        const int c = warp + 16 * q;
        if (c < bw) {
          qp_st_async(rx + (uint32_t)(((par * CS + rank) * BW + c) * sizeof(T)), dots[q], rbar);
          if (rank == owner) qp_st_async(rp + (uint32_t)((par * BW + c) * sizeof(T)), pvs[q], rbar);
        }
      }
    }
    qp_mbar_wait(bar, (uint32_t)(jj >> 1) & 1u);
    // -- reflector scalars, redundantly in every warp: lane c handles column c
    T tot = C::zero(), pv = C::zero();
    if (lane < bw) {
      // fixed-shape pairwise tree over the CS partial sums (same order in every CTA and warp)
      T t4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t4[u] = xbuf[(par * CS + u) * BW + lane];
#pragma unroll
      for (int rk = 4; rk < CS; ++rk) t4[rk & 3] = cx_add<CPLX>(t4[rk & 3], xbuf[(par * CS + rk) * BW + lane]);
      tot = cx_add<CPLX>(cx_add<CPLX>(t4[0], t4[1]), cx_add<CPLX>(t4[2], t4[3]));
      pv = pbuf[par * BW + lane];
    }
    const double ss = __shfl_sync(0xffffffffu, C::re(tot), jj);
    const double ar = __shfl_sync(0xffffffffu, C::re(pv), jj);
    const double ai = CPLX ? __shfl_sync(0xffffffffu, C::im(pv), jj) : 0.0;
    T tau, scale;
    double beta;
    if (ss == 0.0 && ai == 0.0) { tau = C::zero(); scale = C::zero(); beta = ar; }
    else {
      const double nrm = sqrt(ar * ar + ai * ai + ss);
      beta = ar >= 0.0 ? -nrm : nrm;
      const double ib = 1.0 / beta;
      tau = C::make((beta - ar) * ib, -ai * ib);
      const double dr = ar - beta, di = ai, iden = 1.0 / (dr * dr + di * di);
      scale = C::make(dr * iden, -di * iden);
    }
    // lane c > jj: coefficient of the update;  lane c < jj: V^H v_jj entry for the T factor
    T wv = C::zero(), tvec = C::zero();
    if (lane > jj) wv = C::mul(C::conj(tau), cx_add<CPLX>(pv, C::mul(C::conj(scale), tot)));
    else if (lane < jj) tvec = C::mul(C::conj(my_scale), cx_add<CPLX>(C::conj(pv), C::mul(scale, C::conj(tot))));
    else { my_scale = scale; my_tau = tau; my_beta = beta; }
    // -- update this warp's columns to the right of jj
#pragma unroll
    for (int q = 0; q < NCOL; ++q) {
      const int c = warp + 16 * q;
      const double wr = __shfl_sync(0xffffffffu, C::re(wv), c & 31);
      const double wi = CPLX ? __shfl_sync(0xffffffffu, C::im(wv), c & 31) : 0.0;
      if (c > jj && c < bw) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (below[it] || pivot_here[it]) {
            const T v = pivot_here[it] ? C::one() : C::mul(x[it], scale);
            if constexpr (CPLX) {
              a[q][it].x = fma(-wr, v.x, a[q][it].x); a[q][it].x = fma(wi, v.y, a[q][it].x);
              a[q][it].y = fma(-wr, v.y, a[q][it].y); a[q][it].y = fma(-wi, v.x, a[q][it].y);
            } else {
              a[q][it] = fma(-wr, v, a[q][it]);
            }
          }
        }
      }
    }
    // -- next pivot column -> shared broadcast buffer (its owner warp has just updated it)
    if (jj + 1 < bw && warp == ((jj + 1) & 15)) {
      const int qn = (jj + 1) >> 4;
#pragma unroll
      for (int q = 0; q < NCOL; ++q)
        if (q == qn) {
#pragma unroll
          for (int it = 0; it < NIT; ++it) xcol[lane + 32 * it] = a[q][it];
        }
    }
    // -- Gram entries V^H v_jj for the T factor (built after the loop)
    if (rank == 0 && warp == 0) {
      if (lane < jj) Gs[lane * BW + jj] = tvec;
      else if (lane == jj) Ts[jj * BW + jj] = tau;
    }
    __syncthreads();
  }
  // -- store: R part (beta on the diagonal) and the scaled reflectors
#pragma unroll
  for (int q = 0; q < NCOL; ++q) {
    const int c = warp + 16 * q;
    const double sr = __shfl_sync(0xffffffffu, C::re(my_scale), c & 31);
    const double si = CPLX ? __shfl_sync(0xffffffffu, C::im(my_scale), c & 31) : 0.0;
    const double bc = __shfl_sync(0xffffffffu, my_beta, c & 31);
    if (c < bw) {
      const int piv = j0 + c;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int i = lane + 32 * it, r = r0 + i;
        if (i < nloc) {
          const T val = a[q][it];
          At[(long)piv * ldt + r] = r == piv ? C::make(bc, 0.0) : val;
          V[(long)piv * ldt + r] = r < piv ? C::zero() : (r == piv ? C::one() : C::mul(val, C::make(sr, si)));
        }
      }
    }
  }
  if (rank == 0) {
    if (warp == 0 && lane < bw) { tau_out[j0 + lane] = my_tau; rdiag[j0 + lane] = my_beta; }
    // T of H_0 ... H_{bw-1} = I - V T V^H by recursive doubling: for adjacent diagonal blocks A, B of
    // size sz,  T_AB = -T_AA (V_A^H V_B) T_BB  (LAPACK larft's recurrence, log2(BW) levels)
    for (int sz = 1; sz < BW; sz <<= 1) {
      __syncthreads();
      // W = G_AB T_BB
      for (int e = tid; e < (BW / 2) * sz; e += 512) {
        const int pair = e / (sz * sz), rem = e % (sz * sz);
        const int i = pair * 2 * sz + rem / sz, j = pair * 2 * sz + sz + rem % sz;
        const int b0 = pair * 2 * sz + sz;
        T acc = C::zero();
        for (int l = b0; l <= j; ++l) acc = cx_add<CPLX>(acc, C::mul(Gs[i * BW + l], Ts[l * BW + j]));
        Ws[i * BW + j] = acc;
      }
      __syncthreads();
      for (int e = tid; e < (BW / 2) * sz; e += 512) {
        const int pair = e / (sz * sz), rem = e % (sz * sz);
        const int i = pair * 2 * sz + rem / sz, j = pair * 2 * sz + sz + rem % sz;
        const int a1 = pair * 2 * sz + sz;
        T acc = C::zero();
        for (int l = i; l < a1; ++l) acc = cx_add<CPLX>(acc, C::mul(Ts[i * BW + l], Ws[l * BW + j]));
        Ts[i * BW + j] = C::make(-C::re(acc), -C::im(acc));
      }
    }
    __syncthreads();
    for (int i = tid; i < QP_B * QP_B; i += 512) {
      const int ti = i / QP_B, tl = i % QP_B;
      Tout[i] = (ti < BW && tl < BW) ? Ts[ti * BW + tl] : C::zero();
    }
  }
  cluster.sync();      // no CTA may exit while a sibling can still write into its shared memory
}

// ---- 2a. Ypart[chunk][c][i] = sum_{r in chunk} conj(V[j0+i][r]) X[c][r] --------------------------
// grid = (row chunks, column groups), 256 threads, 2 x 2 outputs per thread.
template <bool CPLX>
__global__ void __launch_bounds__(256)
wy_dots_kernel(const typename Cx<CPLX>::T* __restrict__ X, long ldx, int c_begin, int c_end, int m, int j0,
               int bw, const typename Cx<CPLX>::T* __restrict__ V, long ldv,
               typename Cx<CPLX>::T* __restrict__ Ypart, int ncols_pad) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  extern __shared__ __align__(16) unsigned char wy_smem_raw[];
  constexpr int LDS = QP_B + 1;
  T* Vs = reinterpret_cast<T*>(wy_smem_raw);           // [QP_RCH][LDS]  (row, reflector)
  T* Xs = Vs + QP_RCH * LDS;                            // [QP_RCH][LDS]  (row, column)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rbase = j0 + blockIdx.x * QP_RCH;
  const int cbase = c_begin + blockIdx.y * QP_CGW;
  for (int i = warp; i < QP_B; i += 8) {
    const T* vsrc = V + (long)(j0 + i) * ldv;
    const int c = cbase + i;
    const T* xsrc = X + (long)c * ldx;
    for (int rr = lane; rr < QP_RCH; rr += 32) {
      const int r = rbase + rr;
      Vs[rr * LDS + i] = (i < bw && r < m) ? vsrc[r] : C::zero();
      Xs[rr * LDS + i] = (c < c_end && r < m) ? xsrc[r] : C::zero();
    }
  }
  __syncthreads();
  const int ti = tid & 15, tc = tid >> 4;
  double a00r = 0, a00i = 0, a01r = 0, a01i = 0, a10r = 0, a10i = 0, a11r = 0, a11i = 0;
#pragma unroll 4
  for (int rr = 0; rr < QP_RCH; ++rr) {
    const T v0 = Vs[rr * LDS + ti], v1 = Vs[rr * LDS + ti + 16];
    const T x0 = Xs[rr * LDS + tc], x1 = Xs[rr * LDS + tc + 16];
    if constexpr (CPLX) {
      // conj(v) * x accumulated with four FMAs per product
      a00r = fma(v0.x, x0.x, a00r); a00r = fma(v0.y, x0.y, a00r); a00i = fma(v0.x, x0.y, a00i); a00i = fma(-v0.y, x0.x, a00i);
      a01r = fma(v0.x, x1.x, a01r); a01r = fma(v0.y, x1.y, a01r); a01i = fma(v0.x, x1.y, a01i); a01i = fma(-v0.y, x1.x, a01i);
      a10r = fma(v1.x, x0.x, a10r); a10r = fma(v1.y, x0.y, a10r); a10i = fma(v1.x, x0.y, a10i); a10i = fma(-v1.y, x0.x, a10i);
      a11r = fma(v1.x, x1.x, a11r); a11r = fma(v1.y, x1.y, a11r); a11i = fma(v1.x, x1.y, a11i); a11i = fma(-v1.y, x1.x, a11i);
    } else {
      a00r = fma(v0, x0, a00r); a01r = fma(v0, x1, a01r); a10r = fma(v1, x0, a10r); a11r = fma(v1, x1, a11r);
    }
  }
  T* out = Ypart + ((long)blockIdx.x * ncols_pad + (long)blockIdx.y * QP_CGW) * QP_B;
  out[(long)tc * QP_B + ti] = C::make(a00r, a00i);
  out[(long)(tc + 16) * QP_B + ti] = C::make(a01r, a01i);
  out[(long)tc * QP_B + ti + 16] = C::make(a10r, a10i);
  out[(long)(tc + 16) * QP_B + ti + 16] = C::make(a11r, a11i);
}

// ---- 2b. X[c][r] -= sum_i V[j0+i][r] z[i][c],  z = T y (trans = 0) or T^H y (trans = 1) ----------
// y = sum over chunks of Ypart (fixed order).  Same grid as wy_dots_kernel.
template <bool CPLX>
__global__ void __launch_bounds__(256)
wy_update_kernel(typename Cx<CPLX>::T* __restrict__ X, long ldx, int c_begin, int c_end, int m, int j0, int bw,
                 const typename Cx<CPLX>::T* __restrict__ V, long ldv,
                 const typename Cx<CPLX>::T* __restrict__ Tmat, int trans,
                 const typename Cx<CPLX>::T* __restrict__ Ypart, int nchunks, int ncols_pad) {
  pdl_wait();
  using C = Cx<CPLX>;
  using T = typename C::T;
  extern __shared__ __align__(16) unsigned char wy_smem_raw[];
  constexpr int LDY = QP_CGW + 1;
  T* Vs = reinterpret_cast<T*>(wy_smem_raw);           // [QP_B][QP_RCH]  (reflector, row)
  T* ys = Vs + QP_B * QP_RCH;                           // [QP_B][LDY]     (reflector, column)
  T* zs = ys + QP_B * LDY;                              // [QP_B][QP_CGW]  (reflector, column)
  T* Tsm = zs + QP_B * QP_CGW;                          // [QP_B][QP_B]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rbase = j0 + blockIdx.x * QP_RCH;
  const int cbase = c_begin + blockIdx.y * QP_CGW;
  // this thread's 16 entries of X (rows lane + 32 k, columns warp + 8 q): issued first so that the
  // loads overlap the reduction of y and the triangular product
  T xv[4][4];
#pragma unroll
  for (int k4 = 0; k4 < QP_RCH / 32; ++k4)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = rbase + lane + 32 * k4, c = cbase + warp + 8 * q;
      xv[k4][q] = (r < m && c < c_end) ? X[(long)c * ldx + r] : C::zero();
    }
  for (int i = warp; i < QP_B; i += 8) {
    const T* vsrc = V + (long)(j0 + i) * ldv;
    for (int rr = lane; rr < QP_RCH; rr += 32) {
      const int r = rbase + rr;
      Vs[i * QP_RCH + rr] = (i < bw && r < m) ? vsrc[r] : C::zero();
    }
  }
  for (int e = tid; e < QP_B * QP_B; e += 256) Tsm[e] = Tmat[e];
  // y = sum over row chunks of the partial products, fixed order, 16 loads in flight
  for (int e = tid; e < QP_CGW * QP_B; e += 256) {
    T acc = C::zero();
    const T* yp = Ypart + ((long)blockIdx.y * QP_CGW) * QP_B + e;
    for (int ch0 = 0; ch0 < nchunks; ch0 += 16) {
      T part[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) part[u] = (ch0 + u < nchunks) ? yp[(long)(ch0 + u) * ncols_pad * QP_B] : C::zero();
#pragma unroll
      for (int u = 0; u < 16; ++u) acc = cx_add<CPLX>(acc, part[u]);
    }
    ys[(e % QP_B) * LDY + e / QP_B] = acc;            // e = column * QP_B + reflector
  }
  __syncthreads();
  for (int e = tid; e < QP_B * QP_CGW; e += 256) {
    const int i = e / QP_CGW, c = e % QP_CGW;
    T acc = C::zero();
    if (!trans) {
      for (int l = i; l < bw; ++l) acc = cx_add<CPLX>(acc, C::mul(Tsm[i * QP_B + l], ys[l * LDY + c]));
    } else {
      for (int l = 0; l <= i && l < bw; ++l) acc = cx_add<CPLX>(acc, C::cmul(Tsm[l * QP_B + i], ys[l * LDY + c]));
    }
    zs[e] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k4 = 0; k4 < QP_RCH / 32; ++k4) {
    const int rr = lane + 32 * k4, r = rbase + rr;
    T acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = xv[k4][q];
#pragma unroll 4
    for (int i = 0; i < QP_B; ++i) {
      const T v = Vs[i * QP_RCH + rr];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const T z = zs[i * QP_CGW + warp + 8 * q];
        if constexpr (CPLX) {
          acc[q].x = fma(-v.x, z.x, acc[q].x); acc[q].x = fma(v.y, z.y, acc[q].x);
          acc[q].y = fma(-v.x, z.y, acc[q].y); acc[q].y = fma(-v.y, z.x, acc[q].y);
        } else {
          acc[q] = fma(-v, z, acc[q]);
        }
      }
    }
    if (r < m) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = cbase + warp + 8 * q;
        if (c < c_end) X[(long)c * ldx + r] = acc[q];
      }
    }
  }
}

template <bool CPLX>
__global__ void qp_identity_rows_kernel(typename Cx<CPLX>::T* Qt, int m, int k, long ldt) {
  pdl_wait();
  using C = Cx<CPLX>;
  const long total = (long)k * m;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / m), r = (int)(i % m);
    Qt[(long)c * ldt + r] = (r == c) ? C::one() : C::zero();
  }
}

// One compact-WY application on columns [c_begin, c_end) of X, rows >= j0.
template <bool CPLX>
static int wy_apply(cudaStream_t st, typename Cx<CPLX>::T* X, long ldx, int c_begin, int c_end, int m, int j0,
                    int bw, const typename Cx<CPLX>::T* V, long ldv, const typename Cx<CPLX>::T* Tmat, int trans,
                    typename Cx<CPLX>::T* Ypart) {
  using T = typename Cx<CPLX>::T;
  const int ncols = c_end - c_begin;
  if (ncols <= 0 || m - j0 <= 0) return 0;
  const int nchunks = (int)ceil_div(m - j0, QP_RCH);
  const int ngroups = (int)ceil_div(ncols, QP_CGW);
  const int ncols_pad = ngroups * QP_CGW;
  dim3 grid((unsigned)nchunks, (unsigned)ngroups);
  const size_t smem_a = sizeof(T) * 2 * QP_RCH * (QP_B + 1);
  const size_t smem_b = sizeof(T) * (QP_B * QP_RCH + QP_B * (QP_CGW + 1) + QP_CGW * QP_B + QP_B * QP_B);
  static bool attr_done[2] = {false, false};
  if (!attr_done[CPLX ? 1 : 0]) {
    RN_CHECK(cudaFuncSetAttribute(wy_dots_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    RN_CHECK(cudaFuncSetAttribute(wy_update_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    attr_done[CPLX ? 1 : 0] = true;
  }
  { RN_LAUNCH(wy_dots_kernel<CPLX>, grid, 256, smem_a, st, X, ldx, c_begin, c_end, m, j0, bw, V, ldv, Ypart, ncols_pad); rn::g_launches++; }
  { RN_LAUNCH(wy_update_kernel<CPLX>, grid, 256, smem_b, st, X, ldx, c_begin, c_end, m, j0, bw, V, ldv, Tmat, trans, Ypart,
                                                      nchunks, ncols_pad); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// Blocked QR of the m x n matrix held column-as-row in At (n rows of length ldt >= m).
// Returns 1 when the shape is outside what the cluster kernel supports (caller falls back).
template <bool CPLX>
int qr_colmajor_panel(cudaStream_t st, int m, int n, typename Cx<CPLX>::T* At, long ldt,
                      typename Cx<CPLX>::T* V, typename Cx<CPLX>::T* tau, double* rdiag,
                      typename Cx<CPLX>::T* Qt) {
  using T = typename Cx<CPLX>::T;
  const int k = m < n ? m : n;
  // register-resident kernel on 16-CTA clusters when the row slice fits (rloc <= 32 * NIT),
  // else the shared-memory kernel on 8-CTA clusters with the widest panel that fits
  static int cs16_ok = -1;
  if (cs16_ok < 0) {
    cs16_ok = 0;
    if (cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 4, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 8, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 16, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 4, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 8, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 16, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      const int big = 96 * 1024;
      cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaFuncSetAttribute(house_panel_reg_kernel<true, 16, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaFuncSetAttribute(house_panel_reg_kernel<false, 16, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(16); q.blockDim = dim3(512); q.dynamicSmemBytes = 72 * 1024;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, house_panel_reg_kernel<true, 16, 8, 2>, &q) == cudaSuccess && nclusters > 0)
        cs16_ok = 1;
    }
    (void)cudaGetLastError();
    if (const char* e = getenv("RN_QR_REG")) { if (atoi(e) == 0) cs16_ok = 0; }
  }
  const int rloc16 = (int)ceil_div(m, 16);
  int reg_cfg = 0;                       // 0: shared-memory kernel; 1..3: register kernel variants
  if (cs16_ok && rloc16 <= 128) reg_cfg = 1;
  else if (cs16_ok && rloc16 <= 256) reg_cfg = 2;
  else if (cs16_ok && rloc16 <= 512) reg_cfg = 3;
  const int rloc0 = (int)ceil_div(m, QP_CS);
  int bw_max = QP_B;
  const size_t smem_budget = 180 * 1024;
  if (reg_cfg == 3) bw_max = 16;
  if (reg_cfg == 0) {
    while (bw_max > 4 && (size_t)bw_max * rloc0 * sizeof(T) > smem_budget) bw_max >>= 1;
    if ((size_t)bw_max * rloc0 * sizeof(T) > smem_budget) return 1;
    static bool attr_done[2] = {false, false};
    if (!attr_done[CPLX ? 1 : 0]) {
      RN_CHECK(cudaFuncSetAttribute(house_panel_cluster_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_budget));
      attr_done[CPLX ? 1 : 0] = true;
    }
  }
  const int npanels = (int)ceil_div(k, bw_max);
  T* Tall = nullptr;
  T* Ypart = nullptr;
  RN_CHECK(cudaMallocAsync((void**)&Tall, sizeof(T) * (size_t)npanels * QP_B * QP_B, st));
  const long ncols_pad_max = ceil_div(n > k ? n : k, QP_CGW) * QP_CGW;
  RN_CHECK(cudaMallocAsync((void**)&Ypart, sizeof(T) * (size_t)ceil_div(m, QP_RCH) * ncols_pad_max * QP_B, st));
  for (int p = 0; p < npanels; ++p) {
    const int j0 = p * bw_max;
    const int bw = (k - j0) < bw_max ? (k - j0) : bw_max;
    const int cs = reg_cfg ? 16 : QP_CS;
    const int rloc = (int)ceil_div(m - j0, cs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(QP_THREADS);
    // register kernel: xcol + xbuf + pbuf + three BW x BW scratch matrices (sized for double2)
    const int nit = reg_cfg == 1 ? 4 : (reg_cfg == 2 ? 8 : 16), bwk = reg_cfg == 3 ? 16 : 32;
    const size_t reg_smem = sizeof(double2) * (size_t)(nit * 32 + 2 * 16 * bwk + 2 * bwk + 3 * bwk * bwk);
    cfg.dynamicSmemBytes = reg_cfg ? reg_smem : sizeof(T) * (size_t)bw * rloc;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    T* Tp = Tall + (size_t)p * QP_B * QP_B;
    if (reg_cfg == 1)
      RN_CHECK(cudaLaunchKernelEx(&cfg, house_panel_reg_kernel<CPLX, 16, 4, 2>, At, m, ldt, j0, bw, rloc, V, tau, rdiag, Tp));
    else if (reg_cfg == 2)
      RN_CHECK(cudaLaunchKernelEx(&cfg, house_panel_reg_kernel<CPLX, 16, 8, 2>, At, m, ldt, j0, bw, rloc, V, tau, rdiag, Tp));
    else if (reg_cfg == 3)
      RN_CHECK(cudaLaunchKernelEx(&cfg, house_panel_reg_kernel<CPLX, 16, 16, 1>, At, m, ldt, j0, bw, rloc, V, tau, rdiag, Tp));
    else
      RN_CHECK(cudaLaunchKernelEx(&cfg, house_panel_cluster_kernel<CPLX>, At, m, ldt, j0, bw, rloc, V, tau, rdiag, Tp));
    rn::g_launches++;
    int err = wy_apply<CPLX>(st, At, ldt, j0 + bw, n, m, j0, bw, V, ldt, Tp, 1, Ypart);
    if (err) return err;
  }
  // Q = H_0 ... H_{k-1} I, panels applied last to first; panel p only touches columns >= j0
  int nbi = (int)ceil_div((long)k * m, 256);
  if (nbi > 1184) nbi = 1184;
  { RN_LAUNCH(qp_identity_rows_kernel<CPLX>, nbi, 256, 0, st, Qt, m, k, ldt); rn::g_launches++; }
  for (int p = npanels - 1; p >= 0; --p) {
    const int j0 = p * bw_max;
    const int bw = (k - j0) < bw_max ? (k - j0) : bw_max;
    int err = wy_apply<CPLX>(st, Qt, ldt, j0, k, m, j0, bw, V, ldt, Tall + (size_t)p * QP_B * QP_B, 0, Ypart);
    if (err) return err;
  }
  RN_LAUNCH_CHECK();
  RN_CHECK(cudaFreeAsync(Tall, st));
  RN_CHECK(cudaFreeAsync(Ypart, st));
  return 0;
}

template int qr_colmajor_panel<false>(cudaStream_t, int, int, double*, long, double*, double*, double*, double*);
template int qr_colmajor_panel<true>(cudaStream_t, int, int, double2*, long, double2*, double2*, double*, double2*);

}  // namespace rn
