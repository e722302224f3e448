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

namespace rn {

constexpr int OZ_BM = 128, OZ_BN = 128, OZ_BK = 128;     // tile: rows, cols, K bytes (= int8 elems)
constexpr int OZ_STAGES = 5;
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
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const double* __restrict__ sA, const double* __restrict__ sB,
                  double* __restrict__ C, int m, int n, long ldc, int rowsA, int rowsB, int kblocks,
                  int nslices, int tiles_m, int tiles_n) {
  extern __shared__ unsigned char oz_smem_raw[];
  const uint32_t raw = smem_u32(oz_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;                       // 1024-B aligned operand ring
  const uint32_t bars = tiles + OZ_STAGES * 2 * OZ_TILE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * OZ_STAGES;
  const uint32_t tfull_bar = bars + 16 * OZ_STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(oz_smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grouped rasterisation: the ~148 tiles in flight form a block of OZ_GROUP_M row tiles by a
  // dozen column tiles, so that the digit slices they share stay L2 resident and every slice is
  // read from HBM about once
  int tm, tn;
  {
    const int tile = blockIdx.x;
    const int per_group = OZ_GROUP_M * tiles_n;
    const int first_m = (tile / per_group) * OZ_GROUP_M;
    const int gsize = (tiles_m - first_m) < OZ_GROUP_M ? (tiles_m - first_m) : OZ_GROUP_M;
    const int in_group = tile % per_group;
    tm = first_m + in_group % gsize;
    tn = in_group / gsize;
  }
  const int row0 = tm * OZ_BM, col0 = tn * OZ_BN;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int g = 0; g < nslices; ++g)
        for (int s = 0; s <= g; ++s) {
          const int t = g - s;
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            const int st = it % OZ_STAGES;
            const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
            mbar_wait(empty_bar + 8 * st, ph ^ 1u);
            mbar_expect_tx(full_bar + 8 * st, 2 * OZ_TILE_BYTES);
            tma_load_2d(tiles + st * 2 * OZ_TILE_BYTES, &tmA, full_bar + 8 * st, kb * OZ_BK, s * rowsA + row0);
            tma_load_2d(tiles + st * 2 * OZ_TILE_BYTES + OZ_TILE_BYTES, &tmB, full_bar + 8 * st, kb * OZ_BK,
                        t * rowsB + col0);
          }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0;
      for (int g = 0; g < nslices; ++g) {
        const int acc = g & 1;
        const uint32_t use = (uint32_t)(g >> 1);
        mbar_wait(tempty_bar + 8 * acc, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * OZ_BN;
        uint32_t accumulate = 0;
        for (int s = 0; s <= g; ++s)
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            const int st = it % OZ_STAGES;
            const uint32_t ph = (uint32_t)(it / OZ_STAGES) & 1u;
            mbar_wait(full_bar + 8 * st, ph);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(tiles + st * 2 * OZ_TILE_BYTES);
            const uint64_t bdesc = make_smem_desc(tiles + st * 2 * OZ_TILE_BYTES + OZ_TILE_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < OZ_BK / 32; ++k4) {
              umma_i8(d_tmem, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), OZ_IDESC, accumulate);
              accumulate = 1;
            }
            umma_commit(empty_bar + 8 * st);       // stage reusable once these MMAs retire
          }
        umma_commit(tfull_bar + 8 * acc);          // level g complete in TMEM
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
      const int acc = g & 1;
      const uint32_t use = (uint32_t)(g >> 1);
      mbar_wait(tfull_bar + 8 * acc, use & 1u);
      tc_fence_after();
      const double wg = scalbn(1.0, -12 - 7 * g);
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_quarter * 32) << 16) + acc * OZ_BN + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sum[c * 32 + i] = fma((double)(int)v[i], wg, sum[c * 32 + i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
    }
    const int grow = row0 + r;
    if (grow < m) {
      const double sa = sA[grow];
      double* crow = C + (long)grow * ldc;
      const int cbase = col0 + half * 64;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const int gc = cbase + i;
        if (gc < n) crow[gc] = sum[i] * sa * sB[gc];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// --------------------------------------------------------------------------------------- split
// One warp per row: row maximum -> power-of-two scale -> S int8 digits per element.
// q layout: [slice][row][Kp] bytes, Kp a multiple of 16 (zero padded).
__global__ void __launch_bounds__(256)
ozaki_split_kernel(const double* __restrict__ X, long ld, int rows, int K, int Kp, int nslices,
                   signed char* __restrict__ q, double* __restrict__ scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const double* x = X + (long)row * ld;
  double mx = 0.0;
  for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(x[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) frexp(mx, &e);                      // mx = f * 2^e, f in [0.5, 1)
  const double sc = scalbn(1.0, e);
  const double inv = scalbn(1.0, -e);
  if (lane == 0) scale[row] = sc;
  const long slice_stride = (long)rows * Kp;
  signed char* qrow = q + (long)row * Kp;
  for (int k0 = lane * 4; k0 < Kp; k0 += 128) {
    double rr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rr[j] = (k0 + j < K) ? x[k0 + j] * inv * 64.0 : 0.0;
    for (int s = 0; s < nslices; ++s) {
      char4 out;
      signed char* o = reinterpret_cast<signed char*>(&out);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double qd = rint(rr[j]);
        o[j] = (signed char)(int)qd;
        rr[j] = (rr[j] - qd) * 128.0;
      }
      *reinterpret_cast<char4*>(qrow + s * slice_stride + k0) = out;
    }
  }
}

// -------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int ozaki_make_map(CUtensorMap* map, const signed char* q, long total_rows, int Kp);

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

int ozaki_make_map(CUtensorMap* map, const signed char* q, long total_rows, int Kp) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp};
  cuuint32_t box[2] = {(cuuint32_t)OZ_BK, (cuuint32_t)OZ_BM};
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
  { ozaki_split_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(X, ld, rows, K, Kp, nslices, q, scale); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

int launch_ozaki_gemm_maps(cudaStream_t st, int m, int n, int K, int nslices, const CUtensorMap* tmA,
                           const double* sA, const CUtensorMap* tmB, const double* sB, double* C, long ldc) {
  if (m <= 0 || n <= 0) return 0;
  if (nslices < 1 || nslices > OZ_MAX_SLICES) return (int)cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    RN_CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    attr_set = true;
  }
  const int Kp = (K + 15) & ~15;
  const int kblocks = (Kp + OZ_BK - 1) / OZ_BK;
  const int tiles_m = (int)ceil_div(m, OZ_BM), tiles_n = (int)ceil_div(n, OZ_BN);
  { ozaki_gemm_kernel<<<(unsigned)(tiles_m * tiles_n), OZ_THREADS, OZ_SMEM, st>>>(
      *tmA, *tmB, sA, sB, C, m, n, ldc, m, n, kblocks, nslices, tiles_m, tiles_n); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

int launch_ozaki_gemm(cudaStream_t st, int m, int n, int K, int nslices, const signed char* qA,
                      const double* sA, const signed char* qB, const double* sB, double* C, long ldc) {
  if (m <= 0 || n <= 0) return 0;
  const int Kp = (K + 15) & ~15;
  CUtensorMap tmA, tmB;
  int err = ozaki_make_map(&tmA, qA, (long)nslices * m, Kp);
  if (err) return err;
  err = ozaki_make_map(&tmB, qB, (long)nslices * n, Kp);
  if (err) return err;
  return launch_ozaki_gemm_maps(st, m, n, K, nslices, &tmA, sA, &tmB, sB, C, ldc);
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
