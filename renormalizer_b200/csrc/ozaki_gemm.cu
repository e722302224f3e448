// FP64-accurate GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM accumulators,
// TMA operand staging): the Ozaki error-free splitting scheme.
//
//   C[i,j] = sum_k A[i,k] B[j,k]        A (m x K), B (n x K), both K-major FP64
//
// 1. split (HBM-bound, this file): every row of A (and of B) is scaled by a power of two
//    sA_i >= max_k |A_ik| and cut into S signed 8-bit digits
//        A_ik = sA_i * sum_s w_s qA^(s)_ik,   w_s = 2^(-6-7s),  |q| <= 64,
//    all operations exact in FP64.
// 2. digit GEMMs (tensor-bound): P_g = sum_{s+t=g} qA^(s) qB^(t)^T for g = 0..S-1, exact in
//    int32 (K (g+1) 2^12 < 2^31), one TMEM accumulator per level g, double buffered.
// 3. epilogue: C_ij = sA_i sB_j sum_g 2^(-12-7g) P_g,ij accumulated in FP64 registers straight
//    out of TMEM (tcgen05.ld), one pass per level, overlapped with the next level's MMAs.
// Dropped digit pairs (s+t >= S) bound the error by ~ K S 2^(-7S) max|A_i| max|B_j|
// (S = 7: 1e-14 K; S = 8: FP64 level).
//
// Kernel anatomy (one 128x128 output tile per CTA, 320 threads):
//   warp 0      TMA producer      cp.async.bulk.tensor.2d, 128B-swizzled 128x128-byte boxes,
//                                 STAGES-deep mbarrier ring
//   warp 1      MMA issuer        one elected lane: 4 x tcgen05.mma (128x128x32, s8*s8+s32) per
//                                 stage, tcgen05.commit frees the stage / publishes the level
//   warps 2-9   epilogue          tcgen05.ld 32x32b, cvt + FMA into 64 FP64 accumulators/thread
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

#include <cuda.h>
#include <map>
#include <vector>
#include <mutex>
#include <type_traits>

namespace rn {

constexpr int OZ_BM = 128, OZ_BN = 128, OZ_BK = 128;     // tile: rows, cols, K bytes (= int8 elems)
constexpr int OZ_STAGES = 5;
constexpr int OZ_ACC = 4;                                  // TMEM accumulators (128 columns each)
constexpr int OZ_THREADS = 320;
constexpr int OZ_TILE_BYTES = OZ_BM * OZ_BK;             // 16 KiB per operand tile
constexpr int OZ_SMEM = OZ_STAGES * 2 * OZ_TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int OZ_MAX_SLICES = 8;
constexpr int OZ_GROUP_M = 12;

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: operands of one 256 x 128 tile live in the shared memory of
// two CTAs of a cluster, the leader issues the MMAs, barriers of the leader are addressed from the
// peer by clearing the peer bit of the shared-memory address.
constexpr uint32_t OZ_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar & OZ_PEER_MASK), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on `bar` of BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {    // arrive on the leader CTA's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & OZ_PEER_MASK) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// s8 x s8 -> s32, M = 128, N = 128, both operands K-major
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
                              ((uint32_t)(OZ_BM >> 4) << 24);
// the same with M = 256 across a CTA pair
constexpr uint32_t OZ_IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
                               ((uint32_t)((2 * OZ_BM) >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------ the GEMM
// CTA2 = false: one CTA per 128 x 128 tile.  CTA2 = true: a cluster of two CTAs per 256 x 128 tile
// (tcgen05 cta_group::2): each CTA stages its own 128 rows of A and HALF of the B tile (tmB then has
// 64-row boxes) and keeps its 128 x 128 accumulators in its own TMEM; the leader CTA issues the
// MMAs for both.  Per MMA a CTA then reads 6 KB instead of 8 KB from shared memory and TMA writes
// 24 KB instead of 32 KB per step -- shared-memory bandwidth is what bounds this kernel.
// "Unit" below = what one scheduling slot computes: a tile (CTA2 = false) or a tile pair.
template <bool CTA2>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const double* __restrict__ sA, const double* __restrict__ sB,
                  double* __restrict__ C, int m, int n, long ldc, int rowsA, int rowsB, int kblocks,
                  int nslices, int tiles_m, int tiles_n, int n_full, int ksplit_tail, int kb_per_split,
                  double* __restrict__ partial, int* __restrict__ counters,
                  const double* __restrict__ dotv, double* __restrict__ dot_partial) {
  pdl_trigger();
  extern __shared__ unsigned char oz_smem_raw[];
  const uint32_t raw = smem_u32(oz_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;                       // 1024-B aligned operand ring
  const uint32_t bars = tiles + OZ_STAGES * 2 * OZ_TILE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * OZ_STAGES;
  const uint32_t tfull_bar = bars + 16 * OZ_STAGES, tempty_bar = tfull_bar + 8 * OZ_ACC;
  const uint32_t tmem_slot = tempty_bar + 8 * OZ_ACC;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(oz_smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grouped rasterisation: the ~148 tiles in flight form a block of OZ_GROUP_M row tiles by a
  // dozen column tiles, so that the digit slices they share stay L2 resident and every slice is
  // read from HBM about once
  // CTAs [0, n_full) own one whole tile each; the remaining tiles (all of them when n_full == 0)
  // are split along K into ksplit_tail CTAs, so that a ragged last wave finishes in a fraction of
  // a tile time.  The last CTA of a split tile to finish sums the partial tiles in split order
  // (deterministic) and writes C.
  uint32_t rank = 0;                               // CTA rank in the pair
  if constexpr (CTA2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int unit_idx = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  int unit, split, ksplit;
  if (unit_idx < n_full) { unit = unit_idx; split = 0; ksplit = 1; }
  else {
    const int t = unit_idx - n_full;
    unit = n_full + t / ksplit_tail; split = t % ksplit_tail; ksplit = ksplit_tail;
  }
  // per-CTA 128 x 128 tile ids (partial tiles, counters, dot partials)
  const int tile_id = CTA2 ? 2 * unit + (int)rank : unit;
  const int ptile = CTA2 ? 2 * (unit - n_full) + (int)rank : unit - n_full;
  int tm, tn;
  {
    const int tile = unit;
    constexpr int GM = CTA2 ? OZ_GROUP_M / 2 : OZ_GROUP_M;   // tiles_m counts units (256 rows when CTA2)
    const int per_group = GM * tiles_n;
    const int first_m = (tile / per_group) * GM;
    const int gsize = (tiles_m - first_m) < GM ? (tiles_m - first_m) : GM;
    const int in_group = tile % per_group;
    tm = first_m + in_group % gsize;
    tn = in_group / gsize;
  }
  const int row0 = CTA2 ? tm * 2 * OZ_BM + (int)rank * OZ_BM : tm * OZ_BM, col0 = tn * OZ_BN;
  const int bcol0 = CTA2 ? col0 + (int)rank * (OZ_BN / 2) : col0;       // first B row this CTA stages
  const int kb0 = ksplit > 1 ? split * kb_per_split : 0;
  const int kb1 = ksplit > 1 ? ((kb0 + kb_per_split) < kblocks ? (kb0 + kb_per_split) : kblocks) : kblocks;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int a = 0; a < OZ_ACC; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, CTA2 ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(OZ_ACC * OZ_BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(OZ_ACC * OZ_BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();          // both CTAs' barriers exist before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // barriers and TMEM are set up while the previous kernel drains; its results are needed from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Levels are processed in PAIRS (lo, hi = lo + 1) so that operand tiles are shared between digit
  // products: step s of a pair loads A-slice s and B-slice hi - s and feeds TWO products,
  //   A_s x B_(hi-s) -> level hi      and      A_(s-1) x B_(hi-s) -> level lo,
  // the second one re-using the A tile of the previous step.  Per K block that is hi + 1 tile
  // pairs for 2 hi + 1 products instead of one pair per product: L2 -> shared-memory traffic
  // (the bound of this kernel) drops from 56 to 32 tiles per K block at 7 digits.
  // With an odd digit count level 0 runs alone first.
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int g = (nslices & 1) ? -1 : 0; g < nslices; g += 2) {
        const int hi = g + 1;
        for (int kb = kb0; kb < kb1; ++kb)
          for (int s = 0; s <= hi; ++s, ++it) {
            const int st = it % OZ_STAGES;
            const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
            mbar_wait(empty_bar + 8 * st, ph ^ 1u);
            if constexpr (CTA2) {
              // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
              if (rank == 0) mbar_expect_tx(full_bar + 8 * st, 2 * (OZ_TILE_BYTES + OZ_TILE_BYTES / 2));
              tma_load_2d_pair(tiles + st * 2 * OZ_TILE_BYTES, &tmA, full_bar + 8 * st, kb * OZ_BK, s * rowsA + row0);
              tma_load_2d_pair(tiles + st * 2 * OZ_TILE_BYTES + OZ_TILE_BYTES, &tmB, full_bar + 8 * st, kb * OZ_BK,
                               (hi - s) * rowsB + bcol0);
            } else {
              mbar_expect_tx(full_bar + 8 * st, 2 * OZ_TILE_BYTES);
              tma_load_2d(tiles + st * 2 * OZ_TILE_BYTES, &tmA, full_bar + 8 * st, kb * OZ_BK, s * rowsA + row0);
              tma_load_2d(tiles + st * 2 * OZ_TILE_BYTES + OZ_TILE_BYTES, &tmB, full_bar + 8 * st, kb * OZ_BK,
                          (hi - s) * rowsB + col0);
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the leader CTA only when paired) =====
    if (lane == 0 && (!CTA2 || rank == 0)) {
      int it = 0;
      for (int g = (nslices & 1) ? -1 : 0; g < nslices; g += 2) {
        const int lo = g, hi = g + 1;
        if (lo >= 0) mbar_wait(tempty_bar + 8 * (lo & 3), (((uint32_t)lo >> 2) & 1u) ^ 1u);
        mbar_wait(tempty_bar + 8 * (hi & 3), (((uint32_t)hi >> 2) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_lo = tmem_base + (uint32_t)((lo & 3) * OZ_BN), d_hi = tmem_base + (uint32_t)((hi & 3) * OZ_BN);
        uint32_t acc_lo = 0, acc_hi = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          int prev = -1;
          for (int s = 0; s <= hi; ++s, ++it) {
            const int st = it % OZ_STAGES;
            const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
            mbar_wait(full_bar + 8 * st, ph);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(tiles + st * 2 * OZ_TILE_BYTES);
            const uint64_t bdesc = make_smem_desc(tiles + st * 2 * OZ_TILE_BYTES + OZ_TILE_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < OZ_BK / 32; ++k4) {
              if constexpr (CTA2) umma_i8_pair(d_hi, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC2, acc_hi);
              else umma_i8(d_hi, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC, acc_hi);
              acc_hi = 1;
            }
            if (s >= 1) {
              const uint64_t pdesc = make_smem_desc(tiles + prev * 2 * OZ_TILE_BYTES);
#pragma unroll
              for (int k4 = 0; k4 < OZ_BK / 32; ++k4) {
                if constexpr (CTA2) umma_i8_pair(d_lo, pdesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC2, acc_lo);
                else umma_i8(d_lo, pdesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC, acc_lo);
                acc_lo = 1;
              }
              // previous step's tiles are free (in both CTAs) once these retire
              if constexpr (CTA2) umma_commit_pair(empty_bar + 8 * prev); else umma_commit(empty_bar + 8 * prev);
            }
            prev = st;
          }
          if constexpr (CTA2) umma_commit_pair(empty_bar + 8 * prev); else umma_commit(empty_bar + 8 * prev);
        }
        // levels complete in TMEM (of both CTAs)
        if constexpr (CTA2) {
          if (lo >= 0) umma_commit_pair(tfull_bar + 8 * (lo & 3));
          umma_commit_pair(tfull_bar + 8 * (hi & 3));
        } else {
          if (lo >= 0) umma_commit(tfull_bar + 8 * (lo & 3));
          umma_commit(tfull_bar + 8 * (hi & 3));
        }
      }
    }
  } else {
    // ===== epilogue: 8 warps, thread = (row, 64-column half) =====
    const int ew = warp - 2;
    const int lane_quarter = warp & 3;             // TMEM lanes this warp may touch: 32*(warp%4)..
    const int half = ew >> 2;
    const int r = lane_quarter * 32 + lane;        // row inside the tile
    double sum[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) sum[i] = 0.0;
    for (int g = 0; g < nslices; ++g) {
      const int acc = g & 3;
      const uint32_t use = (uint32_t)(g >> 2);
      mbar_wait(tfull_bar + 8 * acc, use & 1u);
      tc_fence_after();
      const double wg = scalbn(1.0, -12 - 7 * g);
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_quarter * 32) << 16) + acc * OZ_BN + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // int32 -> double without the (quarter-rate) I2F: 2^52 + 2^31 + v is exact in the
          // mantissa of a double whose high word is 0x43300000; one subtraction recovers v
          const double x = __hiloint2double(0x43300000, (int)(v[i] ^ 0x80000000u)) - 4503601774854144.0;
          sum[c * 32 + i] = fma(x, wg, sum[c * 32 + i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(tempty_bar + 8 * acc); else mbar_arrive(tempty_bar + 8 * acc); }
    }
    if (ksplit > 1) {
      // partial tiles are private to this kernel: stored thread-major so that every store / load
      // instruction of a warp covers 512 contiguous bytes
      // partial tiles / counters exist for split units only (ptile)
      double2* mine = reinterpret_cast<double2*>(partial + ((long)ptile * ksplit + split) * (OZ_BM * OZ_BN)) +
                      (half * 32) * OZ_BM + r;
#pragma unroll
      for (int i = 0; i < 32; ++i) mine[i * OZ_BM] = make_double2(sum[2 * i], sum[2 * i + 1]);
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ew == 0 && lane == 0) {
        const int old = atomicAdd(counters + ptile, 1);
        const int last = old == ksplit - 1;
        if (last) counters[ptile] = 0;            // self-resetting: ready for the next launch
        *tmem_slot_ptr = (uint32_t)last;          // tmem_base was read by every thread long ago
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool last = *tmem_slot_ptr != 0u;
      if (last) {
        __threadfence();
        const double2* p0 = reinterpret_cast<const double2*>(partial + (long)ptile * ksplit * (OZ_BM * OZ_BN)) +
                            (half * 32) * OZ_BM + r;
#pragma unroll
        for (int i = 0; i < 64; ++i) sum[i] = 0.0;
        for (int sp = 0; sp < ksplit; ++sp) {
          const double2* ps = p0 + (long)sp * (OZ_BM * OZ_BN / 2);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const double2 v = __ldcg(ps + i * OZ_BM);
            sum[2 * i] += v.x; sum[2 * i + 1] += v.y;
          }
        }
      }
      if (!last) goto oz_epilogue_done;
    }
    {
      // C tile through shared memory (the operand ring is idle now): thread-per-row results are
      // transposed so that every global store of a warp writes 256 contiguous bytes of one row
      double* stage = reinterpret_cast<double*>(oz_smem_raw + (tiles - raw));     // [128 cols][129]
      const int grow = row0 + r;
      const double sa = grow < m ? sA[grow] : 0.0;
#pragma unroll
      for (int i = 0; i < 64; ++i) stage[(half * 64 + i) * (OZ_BM + 1) + r] = sum[i] * sa;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      double sb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) sb[j] = (col0 + lane + 32 * j < n) ? sB[col0 + lane + 32 * j] : 0.0;
      // optional fused inner product with a vector laid out like C (the Lanczos alpha = <v_j, H v_j>):
      // one partial per tile, summed in a fixed order -> deterministic
      double dacc = 0.0;
      for (int rr = ew; rr < OZ_BM; rr += 8) {
        const int gr = row0 + rr;
        if (gr >= m) break;
        double* crow = C + (long)gr * ldc + col0;
        const double* vrow = dotv ? dotv + (long)gr * ldc + col0 : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = lane + 32 * j;
          if (col0 + col < n) {
            const double val = stage[col * (OZ_BM + 1) + rr] * sb[j];
            crow[col] = val;
            if (vrow) dacc = fma(val, vrow[col], dacc);
          }
        }
      }
      if (dotv) {
        dacc = warp_sum(dacc);
        asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone is done reading the stage
        if (lane == 0) stage[ew] = dacc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
          double t = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) t += stage[w8];
          dot_partial[2 * tile_id] = t;
          dot_partial[2 * tile_id + 1] = 0.0;
        }
      }
    }
  oz_epilogue_done:;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();          // the peer may still be arriving on / reading from this CTA
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(OZ_ACC * OZ_BN));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(OZ_ACC * OZ_BN));
  }
}

// ------------------------------------------------------------------------------ int8 peak probe
// Dense int8 tensor-pipe rate of this part, measured the way the digit GEMM uses it: one elected
// thread per CTA issues back-to-back tcgen05.mma kind::i8 (128x128x32) on two resident shared-memory
// tiles (no TMA traffic, no epilogue), four TMEM accumulators round-robin.  bench.py's roofline
// denominator (profiles/r02_int8_peak.json) comes from this kernel instead of "2 x bf16".
__global__ void __launch_bounds__(128, 1)
int8_peak_kernel(int iters) {
  extern __shared__ unsigned char pk_smem_raw[];
  const uint32_t raw = smem_u32(pk_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t bar = tiles + 2 * OZ_TILE_BYTES, tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(pk_smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * OZ_TILE_BYTES / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(pk_smem_raw + (tiles - raw))[i] = 0x01010101u * (uint32_t)(i & 3);
  if (warp == 1 && lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(OZ_ACC * OZ_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy tile writes -> async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 1 && lane == 0) {
    const uint64_t adesc = make_smem_desc(tiles), bdesc = make_smem_desc(tiles + OZ_TILE_BYTES);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (uint32_t)((it & 3) * OZ_BN);
#pragma unroll
      for (int k4 = 0; k4 < OZ_BK / 32; ++k4)
        umma_i8(d, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC, it >= 4 || k4 > 0 ? 1u : 0u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(OZ_ACC * OZ_BN));
  }
}

// --------------------------------------------------------------------------------------- split
// Digit extraction shared by the split kernels: v (already divided by the row scale and
// multiplied by 64) -> S signed digits, most significant first.  All steps are exact in FP64.
__device__ __forceinline__ void oz_digits(double rr, int nslices, signed char (&d)[OZ_MAX_SLICES]) {
#pragma unroll
  for (int s = 0; s < OZ_MAX_SLICES; ++s) {
    if (s < nslices) {
      const double qd = rint(rr);
      d[s] = (signed char)(int)qd;
      rr = (rr - qd) * 128.0;
    }
  }
}

__device__ __forceinline__ void oz_scale_of(double mx, double& sc, double& inv64) {
  int e = 0;
  if (mx > 0.0) frexp(mx, &e);                      // mx = f * 2^e, f in [0.5, 1)
  sc = scalbn(1.0, e);
  inv64 = scalbn(1.0, 6 - e);
}

// One warp per source row, K contiguous.
//   FORM 0: real row of K doubles -> one digit row.
//   FORM 1: complex row of K/2 interleaved elements -> the two rows (2r, 2r+1) of its 2x2 real
//           representation ("B-form" of pack.cu):  (re, -im | im, re), with conj_left
//           (re, +im | im, -re).  K counts doubles.
// q layout: [slice][row][Kp] bytes, Kp a multiple of 16 (zero padded).
// WPR warps work on one row (1: a warp per row, 8 rows per block; 8: the whole block on one row,
// for operands with few long rows).
template <int FORM, int WPR>
__global__ void __launch_bounds__(256)
ozaki_split_kernel(const double* __restrict__ X, long ld, int rows, int K, int Kp, int nslices,
                   signed char* __restrict__ q, double* __restrict__ scale, int conj_left) {
  pdl_wait();
  __shared__ double smx[8];
  const int warp = threadIdx.x >> 5;
  const int lane = WPR == 1 ? (threadIdx.x & 31) : threadIdx.x;       // position inside the row team
  constexpr int TEAM = 32 * WPR;
  const int row = WPR == 1 ? blockIdx.x * 8 + warp : blockIdx.x;
  if (row >= rows) return;
  const double* x = X + (long)row * ld;
  double mx = 0.0;
  for (int k = lane; k < K; k += TEAM) mx = fmax(mx, fabs(x[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (WPR > 1) {
    if ((threadIdx.x & 31) == 0) smx[warp] = mx;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) mx = fmax(mx, smx[w]);
  }
  double sc, inv64;
  oz_scale_of(mx, sc, inv64);
  const int out_rows = FORM == 1 ? 2 * rows : rows;
  if (lane == 0) {
    if (FORM == 1) { scale[2 * row] = sc; scale[2 * row + 1] = sc; }
    else scale[row] = sc;
  }
  const long slice_stride = (long)out_rows * Kp;
  signed char* qrow = q + (long)(FORM == 1 ? 2 * row : row) * Kp;
  for (int k0 = lane * 4; k0 < Kp; k0 += 4 * TEAM) {
    signed char dg[4][OZ_MAX_SLICES];
#pragma unroll
    for (int j = 0; j < 4; ++j) oz_digits((k0 + j < K) ? x[k0 + j] * inv64 : 0.0, nslices, dg[j]);
#pragma unroll
    for (int s = 0; s < OZ_MAX_SLICES; ++s) {
      if (s >= nslices) break;
      if (FORM == 0) {
        char4 o = make_char4(dg[0][s], dg[1][s], dg[2][s], dg[3][s]);
        *reinterpret_cast<char4*>(qrow + s * slice_stride + k0) = o;
      } else {
        // (re0, im0, re1, im1) -> row 2r: (re0, -+im0, re1, -+im1); row 2r+1: (im0, +-re0, im1, +-re1)
        const signed char sg = conj_left ? 1 : -1;
        char4 o0 = make_char4(dg[0][s], (signed char)(sg * dg[1][s]), dg[2][s], (signed char)(sg * dg[3][s]));
        char4 o1 = make_char4(dg[1][s], (signed char)(-sg * dg[0][s]), dg[3][s], (signed char)(-sg * dg[2][s]));
        *reinterpret_cast<char4*>(qrow + s * slice_stride + k0) = o0;
        *reinterpret_cast<char4*>(qrow + s * slice_stride + Kp + k0) = o1;
      }
    }
  }
}

// Transposing split: logical operand P[r][c] = src[r + c*s_col] (rows contiguous in memory, the
// contraction index strided) -- the centre tensor C[c,(rest,k)] seen as the K-major right operand
// of G1.  Fuses pack.cu's transpose / realification with the digit split: coalesced reads along
// r, digits staged in shared memory, coalesced 64-byte row segments out.
//   CPLX: out rows (2r, 2r+1) = B-form, K = 2*cols;  real: out row r, K = cols.
// grid = (ceil(rows/32), CY): every block scans all of c for the row maxima (L2 resident) and
// emits the digits of its own share of the c range.
constexpr int OZT_RT = 32;        // source rows per block
constexpr int OZT_KB = 64;        // digit bytes per out row and chunk
constexpr int OZT_LD = OZT_KB + 4;

template <bool CPLX>
__global__ void __launch_bounds__(256)
ozaki_split_t_kernel(const double* __restrict__ src, long s_col, int rows, int cols, int Kp, int nslices,
                     signed char* __restrict__ q, double* __restrict__ scale, int chunks_per_block) {
  pdl_wait();
  constexpr int CT = CPLX ? 32 : 64;             // source columns per chunk
  constexpr int OR = CPLX ? 2 * OZT_RT : OZT_RT; // out rows per block
  __shared__ __align__(16) signed char stage[OZ_MAX_SLICES * OR * OZT_LD];
  __shared__ double smax[8][OZT_RT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * OZT_RT + lane;
  const bool rok = r < rows;
  // ---- row maxima over the whole c range
  double mx = 0.0;
  if (rok) {
    for (int c = warp; c < cols; c += 8) {
      if (CPLX) {
        const double2 v = reinterpret_cast<const double2*>(src)[(long)c * s_col + r];
        mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
      } else {
        mx = fmax(mx, fabs(src[(long)c * s_col + r]));
      }
    }
  }
  smax[warp][lane] = mx;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) mx = fmax(mx, smax[w][lane]);
  double sc, inv64;
  oz_scale_of(mx, sc, inv64);
  if (blockIdx.y == 0 && warp == 0 && rok) {
    if (CPLX) { scale[2 * r] = sc; scale[2 * r + 1] = sc; }
    else scale[r] = sc;
  }
  const int out_rows = CPLX ? 2 * rows : rows;
  const long slice_stride = (long)out_rows * Kp;
  const int nchunks = (Kp + OZT_KB - 1) / OZT_KB;
  const int ch0 = blockIdx.y * chunks_per_block;
  const int ch1 = (ch0 + chunks_per_block) < nchunks ? (ch0 + chunks_per_block) : nchunks;
  for (int ch = ch0; ch < ch1; ++ch) {
    const int c0 = ch * CT;
    __syncthreads();                              // stage free again
#pragma unroll
    for (int i = 0; i < CT / 8; ++i) {
      const int cl = warp + 8 * i, c = c0 + cl;
      signed char d0[OZ_MAX_SLICES], d1[OZ_MAX_SLICES];
      if (CPLX) {
        double2 v = make_double2(0.0, 0.0);
        if (rok && c < cols) v = reinterpret_cast<const double2*>(src)[(long)c * s_col + r];
        oz_digits(v.x * inv64, nslices, d0);
        oz_digits(v.y * inv64, nslices, d1);
#pragma unroll
        for (int s = 0; s < OZ_MAX_SLICES; ++s)
          if (s < nslices) {
            signed char* b = stage + (s * OR + 2 * lane) * OZT_LD + 2 * cl;
            *reinterpret_cast<char2*>(b) = make_char2(d0[s], (signed char)(-d1[s]));
            *reinterpret_cast<char2*>(b + OZT_LD) = make_char2(d1[s], d0[s]);
          }
      } else {
        double v = 0.0;
        if (rok && c < cols) v = src[(long)c * s_col + r];
        oz_digits(v * inv64, nslices, d0);
#pragma unroll
        for (int s = 0; s < OZ_MAX_SLICES; ++s)
          if (s < nslices) stage[(s * OR + lane) * OZT_LD + cl] = d0[s];
      }
    }
    __syncthreads();
    // ---- write out: (slice, out row) segments of 64 bytes, 16 lanes x char4 each
    const int kbase = ch * OZT_KB;
    const int seg_total = nslices * OR;
    for (int seg = warp * 2 + (lane >> 4); seg < seg_total; seg += 16) {
      const int s = seg / OR, orow = seg % OR;
      const int grow = blockIdx.x * OR + orow;
      const int kk = (lane & 15) * 4;
      if (grow < out_rows && kbase + kk < Kp) {
        const char4 v = *reinterpret_cast<const char4*>(stage + (s * OR + orow) * OZT_LD + kk);
        *reinterpret_cast<char4*>(q + s * slice_stride + (long)grow * Kp + kbase + kk) = v;
      }
    }
  }
}

// MPO-site application fused with the digit split of its result (the left operand of G3):
//   row (x, d, y1) of  T[x,d,y1,f,y2] = sum_{p,q} W[p,d,q,f] in[x,p,q,(y1,y2)]   (wapply.cu)
// is produced in shared memory by one block, which then scales it and writes its digits; the FP64
// tensor T never touches HBM.  Requires the input's y index to be contiguous.
template <bool CPLX>
__global__ void __launch_bounds__(256)
wapply_split_kernel(WApplyParams p, int Kp, int nslices, signed char* __restrict__ q,
                    double* __restrict__ scale) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  extern __shared__ __align__(16) double ws_row[];
  __shared__ double red[8];
  const int Y1 = p.Y / p.Y2;
  const int row = blockIdx.x;
  const int y1 = row % Y1, xd = row / Y1;
  const int d = xd % p.D, x = xd / p.D;
  const T* in = reinterpret_cast<const T*>(p.in) + (long)x * p.isx + (long)y1 * p.Y2;
  // CSR entries of this d (all f): element offsets and values staged once per block, so the inner
  // loop has no integer division and no dependent index loads
  constexpr int WS_MAXE = 256;
  __shared__ int s_off[WS_MAXE];
  __shared__ double s_val[WS_MAXE];
  const int ebase = p.rowptr[d * p.F], eend = p.rowptr[(d + 1) * p.F];
  const bool staged = (eend - ebase) <= WS_MAXE && (long)p.P * p.isp + (long)p.Q * p.isq < 2147483647L;
  if (staged) {
    for (int e = ebase + threadIdx.x; e < eend; e += 256) {
      const int pq = p.ent_pq[e];
      s_off[e - ebase] = (int)((long)(pq / p.Q) * p.isp + (long)(pq % p.Q) * p.isq);
      s_val[e - ebase] = p.ent_val[e];
    }
    __syncthreads();
  }
  double mx = 0.0;
  for (int f = 0; f < p.F; ++f) {
    const int e0 = p.rowptr[d * p.F + f], e1 = p.rowptr[d * p.F + f + 1];
    for (int y2 = threadIdx.x; y2 < p.Y2; y2 += 256) {
      T acc;
      if constexpr (CPLX) acc = make_double2(0.0, 0.0); else acc = 0.0;
      if (staged) {
        for (int e = e0 - ebase; e < e1 - ebase; ++e) {
          const double w = s_val[e];
          const T v = in[s_off[e] + y2];
          if constexpr (CPLX) { acc.x = fma(w, v.x, acc.x); acc.y = fma(w, v.y, acc.y); }
          else acc = fma(w, v, acc);
        }
      } else {
        for (int e = e0; e < e1; ++e) {
          const int pq = p.ent_pq[e];
          const double w = p.ent_val[e];
          const T v = in[(long)(pq / p.Q) * p.isp + (long)(pq % p.Q) * p.isq + y2];
          if constexpr (CPLX) { acc.x = fma(w, v.x, acc.x); acc.y = fma(w, v.y, acc.y); }
          else acc = fma(w, v, acc);
        }
      }
      if constexpr (CPLX) {
        reinterpret_cast<double2*>(ws_row)[f * p.Y2 + y2] = acc;
        mx = fmax(mx, fmax(fabs(acc.x), fabs(acc.y)));
      } else {
        ws_row[f * p.Y2 + y2] = acc;
        mx = fmax(mx, fabs(acc));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) mx = fmax(mx, red[w]);
  double sc, inv64;
  oz_scale_of(mx, sc, inv64);
  if (threadIdx.x == 0) scale[row] = sc;
  const int K = p.F * p.Y2 * (CPLX ? 2 : 1);
  const long slice_stride = (long)gridDim.x * Kp;
  signed char* qrow = q + (long)row * Kp;
  for (int k0 = threadIdx.x * 4; k0 < Kp; k0 += 1024) {
    signed char dg[4][OZ_MAX_SLICES];
#pragma unroll
    for (int j = 0; j < 4; ++j) oz_digits((k0 + j < K) ? ws_row[k0 + j] * inv64 : 0.0, nslices, dg[j]);
#pragma unroll
    for (int s = 0; s < OZ_MAX_SLICES; ++s) {
      if (s >= nslices) break;
      *reinterpret_cast<char4*>(qrow + s * slice_stride + k0) = make_char4(dg[0][s], dg[1][s], dg[2][s], dg[3][s]);
    }
  }
}

// -------------------------------------------------------------------------------------- host
static int g_oz_sms = -1;

// Optional per-launch timing of the GEMM kernels (bench.py's roofline leg): between
// rn_profile_begin and rn_profile_end every launch of the given kind is bracketed by CUDA events
// on its own stream.  kind 1 = tcgen05 digit GEMM, kind 0 = FP64 DMMA GEMM.
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev[2];
  double flops[2] = {0.0, 0.0};
};
static GemmProfile g_prof;
bool gemm_profile_on() { return g_prof.on; }
static std::mutex g_prof_mu;      // block SVDs of one bond run on several host threads / streams
cudaEvent_t gemm_profile_begin(cudaStream_t st) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  cudaEventRecord(e, st);
  return e;
}
void gemm_profile_end(cudaStream_t st, int kind, cudaEvent_t begin, double flops) {
  cudaEvent_t e = nullptr;
  if (begin == nullptr || cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.ev[kind].push_back(begin);
  g_prof.ev[kind].push_back(e);
  g_prof.flops[kind] += flops;
}

// Split-K scratch (partial tiles + self-resetting tile counters), one per stream: launches on one
// stream are ordered, so consecutive GEMMs can share it.
struct OzScratch {
  double* partial = nullptr;
  size_t partial_bytes = 0;
  int* counters = nullptr;
  int ncounters = 0;
};
static std::mutex g_oz_scratch_mu;
static std::map<cudaStream_t, OzScratch> g_oz_scratch;
static OzScratch& oz_scratch(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_oz_scratch_mu);
  return g_oz_scratch[st];
}
static int g_oz_force_ksplit = 0;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int ozaki_make_map(CUtensorMap* map, const signed char* q, long total_rows, int Kp, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp};
  cuuint32_t box[2] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)q, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[rn_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return (int)cudaErrorInvalidValue;
  }
  return 0;
}

size_t ozaki_split_bytes(int rows, int K, int nslices) {
  const int Kp = (K + 15) & ~15;
  return (size_t)nslices * rows * Kp;
}

int launch_ozaki_split(cudaStream_t st, const double* X, long ld, int rows, int K, int nslices,
                       signed char* q, double* scale) {
  if (rows <= 0) return 0;
  const int Kp = (K + 15) & ~15;
  if (rows >= 2048 || K < 512)
    { RN_LAUNCH((ozaki_split_kernel<0, 1>), (unsigned)ceil_div(rows, 8), 256, 0, st, X, ld, rows, K, Kp, nslices, q, scale, 0); rn::g_launches++; }
  else
    { RN_LAUNCH((ozaki_split_kernel<0, 8>), (unsigned)rows, 256, 0, st, X, ld, rows, K, Kp, nslices, q, scale, 0); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// Complex source rows (crows x ccols interleaved elements, leading dimension ld in DOUBLES) ->
// digits of the 2*crows x 2*ccols B-form operand.
int launch_ozaki_split_bform(cudaStream_t st, const double* X, long ld, int crows, int ccols, int nslices,
                             int conj_left, signed char* q, double* scale) {
  if (crows <= 0) return 0;
  const int K = 2 * ccols, Kp = (K + 15) & ~15;
  if (crows >= 2048 || K < 512)
    { RN_LAUNCH((ozaki_split_kernel<1, 1>), (unsigned)ceil_div(crows, 8), 256, 0, st, X, ld, crows, K, Kp, nslices, q, scale, conj_left); rn::g_launches++; }
  else
    { RN_LAUNCH((ozaki_split_kernel<1, 8>), (unsigned)crows, 256, 0, st, X, ld, crows, K, Kp, nslices, q, scale, conj_left); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// Transposed view: logical P[r][c] = src[r + c*s_col] (elements), see ozaki_split_t_kernel.
int launch_ozaki_split_t(cudaStream_t st, int cplx, const void* src, long s_col, int rows, int cols,
                         int nslices, signed char* q, double* scale) {
  if (rows <= 0) return 0;
  const int K = cplx ? 2 * cols : cols, Kp = (K + 15) & ~15;
  const int nchunks = (Kp + OZT_KB - 1) / OZT_KB;
  const int bx = (int)ceil_div(rows, OZT_RT);
  int by = (int)ceil_div(296, bx);
  if (by > nchunks) by = nchunks;
  if (by < 1) by = 1;
  const int cpb = (nchunks + by - 1) / by;
  by = (nchunks + cpb - 1) / cpb;
  dim3 grid((unsigned)bx, (unsigned)by);
  if (cplx) { RN_LAUNCH(ozaki_split_t_kernel<true>, grid, 256, 0, st, (const double*)src, s_col, rows, cols, Kp, nslices, q, scale, cpb); rn::g_launches++; }
  else { RN_LAUNCH(ozaki_split_t_kernel<false>, grid, 256, 0, st, (const double*)src, s_col, rows, cols, Kp, nslices, q, scale, cpb); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// Fused MPO application + split (see wapply_split_kernel).  Returns 1 when the row does not fit in
// shared memory or the layout is not supported (caller falls back to wapply + split), 0 on
// success, a CUDA error otherwise.
int launch_wapply_split(cudaStream_t st, int cplx, const WApplyParams& p, int nslices, signed char* q,
                        double* scale) {
  if (p.isy != 1 || p.Y2 <= 0 || p.Y % p.Y2 != 0) return 1;
  const int es = cplx ? 2 : 1;
  const long K = (long)p.F * p.Y2 * es;
  const size_t smem = (size_t)((K + 3) & ~3L) * sizeof(double);
  if (smem > 160 * 1024) return 1;
  const long rows = (long)p.X * p.D * (p.Y / p.Y2);
  if (rows <= 0) return 0;
  const int Kp = (int)((K + 15) & ~15L);
  static size_t max_set[2] = {0, 0};
  if (smem + 1024 > 48 * 1024 && smem > max_set[cplx]) {
    if (cplx) RN_CHECK(cudaFuncSetAttribute(wapply_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    else RN_CHECK(cudaFuncSetAttribute(wapply_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    max_set[cplx] = 160 * 1024;
  }
  if (cplx) { RN_LAUNCH(wapply_split_kernel<true>, (unsigned)rows, 256, smem, st, p, Kp, nslices, q, scale); rn::g_launches++; }
  else { RN_LAUNCH(wapply_split_kernel<false>, (unsigned)rows, 256, smem, st, p, Kp, nslices, q, scale); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// CTA-pair (cta_group::2) kernel for m > 128: validated, but measured no faster than one CTA per
// tile on B200 (M=1024: 90.6 vs 95.6 TFLOP/s FP64-equivalent, the digit GEMMs already sit at the
// sustained int8 rate under the power cap; M=256: 52 vs 55) -- opt-in with RN_OZ_CTA2=1.
static int g_oz_cta2 = -1;
static bool oz_use_pair(int m) {
  if (g_oz_cta2 < 0) { const char* e = getenv("RN_OZ_CTA2"); g_oz_cta2 = (e && e[0] == '1') ? 1 : 0; }
  return g_oz_cta2 && m > OZ_BM;
}

// Number of 128 x 128 CTA tiles (= inner-product partials of the fused dot) of an m x n product.
int ozaki_gemm_tiles(int m, int n) {
  const long tn = ceil_div(n, OZ_BN);
  return (int)(oz_use_pair(m) ? 2 * ceil_div(m, 2 * OZ_BM) * tn : ceil_div(m, OZ_BM) * tn);
}

int launch_ozaki_gemm_maps(cudaStream_t st, int m, int n, int K, int nslices, const CUtensorMap* tmA,
                           const double* sA, const CUtensorMap* tmB, const CUtensorMap* tmB64, const double* sB,
                           double* C, long ldc, const double* dotv, double* dot_partial) {
  if (m <= 0 || n <= 0) return 0;
  if (nslices < 1 || nslices > OZ_MAX_SLICES) return (int)cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    RN_CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    RN_CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    attr_set = true;
  }

  const bool pair = tmB64 != nullptr && oz_use_pair(m);
  const int Kp = (K + 15) & ~15;
  const int kblocks = (Kp + OZ_BK - 1) / OZ_BK;
  // scheduling units: tiles, or 256 x 128 tile pairs run by two CTAs on neighbouring SMs
  const int tiles_m = (int)ceil_div(m, pair ? 2 * OZ_BM : OZ_BM), tiles_n = (int)ceil_div(n, OZ_BN);
  const int tiles = tiles_m * tiles_n;
  const int per_unit = pair ? 2 : 1;
  // split-K when the unit count leaves SMs idle: pick the K partition with the smallest modelled
  // time  waves * (K blocks per CTA * products * 256 clk + fixed CTA cost) + reduction
  int ksplit = 1, kb_per = kblocks, n_full = 0;
  if (g_oz_sms < 0) {
    int dev = 0;
    RN_CHECK(cudaGetDevice(&dev));
    RN_CHECK(cudaDeviceGetAttribute(&g_oz_sms, cudaDevAttrMultiProcessorCount, dev));
    if (const char* e = getenv("RN_OZ_KSPLIT")) g_oz_force_ksplit = atoi(e);
  }
  const int slots = g_oz_sms / per_unit;
  // a CTA's int32 level accumulators overflow beyond 65536 contraction elements (512 K blocks): longer
  // contractions MUST be split along K, the partial tiles are summed in FP64
  constexpr int OZ_KB_MAX = 65536 / OZ_BK;
  const int min_split = (kblocks + OZ_KB_MAX - 1) / OZ_KB_MAX;
  if (min_split > 16) return (int)cudaErrorInvalidValue;
  {
    const double t_kb = 256.0 * (nslices * (nslices + 1) / 2), t_fixed = 9000.0, t_red = 1200.0;
    double best = 1e300;
    const int smax = kblocks < 16 ? kblocks : 16;
    // (a) every unit split the same way
    for (int s = min_split; s <= smax; ++s) {
      const int per = (kblocks + s - 1) / s;
      const int seff = (kblocks + per - 1) / per;
      if (seff != s) continue;
      // partial tiles stay L2 resident -- unless the split is forced by the accumulator range
      if (min_split == 1 && (double)tiles * per_unit * seff * OZ_BM * OZ_BN * 8.0 > 192e6) continue;
      const long waves = ceil_div((long)tiles * seff, slots);
      const double cost = waves * (per * t_kb + t_fixed) + (seff > 1 ? t_red * seff + 2000.0 : 0.0);
      if (cost < best * 0.97) { best = cost; ksplit = seff; kb_per = per; n_full = 0; }
    }
    // (b) whole waves unsplit, the ragged last wave split along K
    const int tail = tiles % slots, full = tiles - tail;
    if (full > 0 && tail > 0 && min_split == 1) {
      const int st_max = slots / tail < smax ? slots / tail : smax;
      for (int s = 2; s <= st_max; ++s) {
        const int per = (kblocks + s - 1) / s;
        const int seff = (kblocks + per - 1) / per;
        if (seff != s) continue;
        const double cost = (full / slots) * (kblocks * t_kb + t_fixed) + (per * t_kb + t_fixed) + t_red * seff + 2000.0;
        if (cost < best * 0.97) { best = cost; ksplit = seff; kb_per = per; n_full = full; }
      }
    }
    if (g_oz_force_ksplit >= min_split && g_oz_force_ksplit <= kblocks) {
      kb_per = (kblocks + g_oz_force_ksplit - 1) / g_oz_force_ksplit;
      ksplit = (kblocks + kb_per - 1) / kb_per;
      n_full = 0;
    }
  }
  if (ksplit < min_split) {                 // no balanced partition qualified: the plain forced split
    kb_per = (kblocks + min_split - 1) / min_split;
    ksplit = (kblocks + kb_per - 1) / kb_per;
    n_full = 0;
  }
  if (ksplit == 1) n_full = tiles;
  const int split_units = tiles - n_full;
  double* partial = nullptr;
  int* counters = nullptr;
  if (split_units > 0) {
    // per-stream scratch that outlives the launch: the tile counters reset themselves (last CTA),
    // so they are zeroed once, and no allocation / memset node sits between the kernels of a chain
    OzScratch& sc = oz_scratch(st);
    const int split_tiles = split_units * per_unit;
    const size_t pbytes = (size_t)split_tiles * ksplit * OZ_BM * OZ_BN * sizeof(double);
    if (pbytes > sc.partial_bytes) {
      if (sc.partial) RN_CHECK(cudaFreeAsync(sc.partial, st));
      const size_t want = pbytes < (32u << 20) ? (32u << 20) : pbytes;
      RN_CHECK(cudaMallocAsync((void**)&sc.partial, want, st));
      sc.partial_bytes = want;
    }
    if (split_tiles > sc.ncounters) {
      if (sc.counters) RN_CHECK(cudaFreeAsync(sc.counters, st));
      const int want = split_tiles < 4096 ? 4096 : split_tiles;
      RN_CHECK(cudaMallocAsync((void**)&sc.counters, sizeof(int) * (size_t)want, st));
      RN_CHECK(cudaMemsetAsync(sc.counters, 0, sizeof(int) * (size_t)want, st));
      sc.ncounters = want;
    }
    partial = sc.partial;
    counters = sc.counters;
  }
  const unsigned units_launched = (unsigned)(n_full + split_units * ksplit);
  cudaEvent_t prof_begin = g_prof.on ? gemm_profile_begin(st) : nullptr;
  const int ks_arg = ksplit > 1 ? ksplit : 1;
  if (!pair) {
    { RN_LAUNCH(ozaki_gemm_kernel<false>, units_launched, OZ_THREADS, OZ_SMEM, st,
        *tmA, *tmB, sA, sB, C, m, n, ldc, m, n, kblocks, nslices, tiles_m, tiles_n, n_full, ks_arg, kb_per,
        partial, counters, dotv, dot_partial); rn::g_launches++; }
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units_launched); cfg.blockDim = dim3(OZ_THREADS);
    cfg.dynamicSmemBytes = OZ_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    RN_CHECK(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<true>, *tmA, *tmB64, sA, sB, C, m, n, ldc, m, n, kblocks,
                                nslices, tiles_m, tiles_n, n_full, ks_arg, kb_per, partial, counters, dotv, dot_partial));
    rn::g_launches++;
  }
  if (prof_begin) gemm_profile_end(st, 1, prof_begin, 2.0 * m * n * K);
  RN_LAUNCH_CHECK();
  return 0;
}

int launch_ozaki_gemm(cudaStream_t st, int m, int n, int K, int nslices, const signed char* qA,
                      const double* sA, const signed char* qB, const double* sB, double* C, long ldc) {
  if (m <= 0 || n <= 0) return 0;
  const int Kp = (K + 15) & ~15;
  CUtensorMap tmA, tmB;
  CUtensorMap tmB64;
  int err = ozaki_make_map(&tmA, qA, (long)nslices * m, Kp, OZ_BM);
  if (err) return err;
  err = ozaki_make_map(&tmB, qB, (long)nslices * n, Kp, OZ_BN);
  if (err) return err;
  err = ozaki_make_map(&tmB64, qB, (long)nslices * n, Kp, OZ_BN / 2);
  if (err) return err;
  return launch_ozaki_gemm_maps(st, m, n, K, nslices, &tmA, sA, &tmB, &tmB64, sB, C, ldc, nullptr, nullptr);
}

}  // namespace rn

// Convenience entry point: split both operands and multiply (used by the tests and by plans that
// have not cached their splits).
extern "C" int rn_ozaki_gemm_tn(void* stream, int m, int n, int k, const double* A, long lda,
                                const double* B, long ldb, double* C, long ldc, int nslices) {
  using namespace rn;
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0 || n <= 0) return 0;
  signed char *qA = nullptr, *qB = nullptr;
  double *sA = nullptr, *sB = nullptr;
  RN_CHECK(cudaMallocAsync((void**)&qA, ozaki_split_bytes(m, k, nslices) + 16, st));
  RN_CHECK(cudaMallocAsync((void**)&qB, ozaki_split_bytes(n, k, nslices) + 16, st));
  RN_CHECK(cudaMallocAsync((void**)&sA, sizeof(double) * m, st));
  RN_CHECK(cudaMallocAsync((void**)&sB, sizeof(double) * n, st));
  int err = launch_ozaki_split(st, A, lda, m, k, nslices, qA, sA);
  if (err) return err;
  err = launch_ozaki_split(st, B, ldb, n, k, nslices, qB, sB);
  if (err) return err;
  err = launch_ozaki_gemm(st, m, n, k, nslices, qA, sA, qB, sB, C, ldc);
  if (err) return err;
  cudaFreeAsync(qA, st); cudaFreeAsync(qB, st); cudaFreeAsync(sA, st); cudaFreeAsync(sB, st);
  return 0;
}

// Dense int8 tcgen05 rate: `iters` MMAs of 128x128x128 (4 instructions of K = 32) per CTA on one CTA
// per SM; *tops_out = 1e-12 * ops / s of this launch, timed with CUDA events on `stream`.
extern "C" int rn_int8_peak(void* stream, int iters, double* tops_out) {
  using namespace rn;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  RN_CHECK(cudaGetDevice(&dev));
  RN_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = 2 * OZ_TILE_BYTES + 1024 + 64;
  RN_CHECK(cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  RN_CHECK(cudaEventCreate(&e0));
  RN_CHECK(cudaEventCreate(&e1));
  RN_CHECK(cudaEventRecord(e0, st));
  int8_peak_kernel<<<sms, 128, smem, st>>>(iters);
  RN_CHECK(cudaEventRecord(e1, st));
  RN_CHECK(cudaEventSynchronize(e1));
  RN_LAUNCH_CHECK();
  float ms = 0.f;
  RN_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (tops_out) *tops_out = (double)sms * iters * 2.0 * OZ_BM * OZ_BN * OZ_BK / (ms * 1e-3) * 1e-12;
  return 0;
}

// ---- GEMM launch profiling (used by bench.py only) ---------------------------------------------
extern "C" int rn_profile_begin(void) {
  using namespace rn;
  g_prof.on = true;
  for (int k = 0; k < 2; ++k) { g_prof.ev[k].clear(); g_prof.flops[k] = 0.0; }
  return 0;
}

// Totals of the tcgen05 digit GEMM launches when there were any, else of the DMMA launches.
extern "C" int rn_profile_end(double* total_ms, double* total_flops, long* launches) {
  using namespace rn;
  g_prof.on = false;
  RN_CHECK(cudaDeviceSynchronize());
  const int kind = g_prof.ev[1].empty() ? 0 : 1;
  double ms = 0.0;
  for (size_t i = 0; i + 1 < g_prof.ev[kind].size(); i += 2) {
    float t = 0.f;
    RN_CHECK(cudaEventElapsedTime(&t, g_prof.ev[kind][i], g_prof.ev[kind][i + 1]));
    ms += t;
  }
  for (int k = 0; k < 2; ++k) {
    for (cudaEvent_t e : g_prof.ev[k]) cudaEventDestroy(e);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = g_prof.flops[kind];
  if (launches) *launches = (long)(g_prof.ev[kind].size() / 2);
  for (int k = 0; k < 2; ++k) g_prof.ev[k].clear();
  return 0;
}
