// FP64 "TN" GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64), cp.async 3-stage pipeline.
//
//   C[i*ldc + j] (+)= sum_k A[i*lda + k] * B[j*ldb + k]      0<=i<m, 0<=j<n, 0<=k<K
//
// Both operands are K-major.  Every tensor-network contraction of the sweep path is lowered to
// this one shape by the pack kernels (pack.cu): complex operands are "realified" there, so this
// kernel only ever sees real data (a complex (m x k)(k x n) product is a real (m x 2k)(2k x 2n)
// one).  This is the exact-FP64 path; the tcgen05 int8 split path (ozaki_gemm.cu) replaces it for
// large contractions and is validated against it.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

namespace rn {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_BKP = 20, G_STAGES = 3, G_THREADS = 256;
constexpr int G_SMEM_BYTES = G_STAGES * (G_BM + G_BN) * G_BKP * (int)sizeof(double);

// Load a 128 x 16 tile (rows row0.., k range k0..) of a K-major matrix into padded smem.
template <bool VEC16>
__device__ __forceinline__ void g_load_tile(double* s, const double* __restrict__ g, int row0,
                                            int nrows, long ld, int k0, int K, int tid) {
  if (VEC16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * G_THREADS;
      const int r = c >> 3, kc = (c & 7) * 2;
      const int gr = row0 + r, gk = k0 + kc;
      int valid = 0;
      if (gr < nrows) {
        const int rem = K - gk;
        valid = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
      }
      const double* src = valid ? g + (long)gr * ld + gk : g;
      cp_async_zfill<16>(s + r * G_BKP + kc, src, valid);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = tid + i * G_THREADS;
      const int r = c >> 4, kc = c & 15;
      const int gr = row0 + r, gk = k0 + kc;
      const int valid = (gr < nrows && gk < K) ? 8 : 0;
      const double* src = valid ? g + (long)gr * ld + gk : g;
      cp_async_zfill<8>(s + r * G_BKP + kc, src, valid);
    }
  }
}

__device__ __forceinline__ void dmma_8x8x4(double (&d)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d[0]), "+d"(d[1])
      : "d"(a), "d"(b));
}

template <bool VEC16>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tn_f64_kernel(const double* __restrict__ A, const double* __restrict__ B,
                   double* __restrict__ C, int m, int n, int K, long lda, long ldb, long ldc,
                   long sA, long sB, long sC, int accumulate) {
  pdl_wait();
  extern __shared__ __align__(16) double g_smem[];
  double* As = g_smem;
  double* Bs = g_smem + G_STAGES * G_BM * G_BKP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1;  // 4 warps along M, 32 rows each
  const int wn = warp & 1;   // 2 warps along N, 64 cols each
  A += (long)blockIdx.z * sA;
  B += (long)blockIdx.z * sB;
  C += (long)blockIdx.z * sC;
  const int row0 = blockIdx.y * G_BM, col0 = blockIdx.x * G_BN;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + G_BK - 1) / G_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < nk) {
      g_load_tile<VEC16>(As + s * G_BM * G_BKP, A, row0, m, lda, s * G_BK, K, tid);
      g_load_tile<VEC16>(Bs + s * G_BN * G_BKP, B, col0, n, ldb, s * G_BK, K, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    const int nt = kt + G_STAGES - 1;
    if (nt < nk) {
      const int st = nt % G_STAGES;
      g_load_tile<VEC16>(As + st * G_BM * G_BKP, A, row0, m, lda, nt * G_BK, K, tid);
      g_load_tile<VEC16>(Bs + st * G_BN * G_BKP, B, col0, n, ldb, nt * G_BK, K, tid);
    }
    cp_async_commit();
    const double* as = As + (kt % G_STAGES) * G_BM * G_BKP + (wm * 32 + (lane >> 2)) * G_BKP + (lane & 3);
    const double* bs = Bs + (kt % G_STAGES) * G_BN * G_BKP + (wn * 64 + (lane >> 2)) * G_BKP + (lane & 3);
#pragma unroll
    for (int ks = 0; ks < G_BK / 4; ++ks) {
      double a[4], b[8];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[mi] = as[mi * 8 * G_BKP + ks * 4];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) b[ni] = bs[ni * 8 * G_BKP + ks * 4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) dmma_8x8x4(acc[mi][ni], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each thread owns (row, 2 consecutive cols) of every 8x8 mma tile
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int r = row0 + wm * 32 + mi * 8 + (lane >> 2);
    if (r >= m) continue;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int c = col0 + wn * 64 + ni * 8 + 2 * (lane & 3);
      double* p = C + (long)r * ldc + c;
      if (c < n) p[0] = accumulate ? p[0] + acc[mi][ni][0] : acc[mi][ni][0];
      if (c + 1 < n) p[1] = accumulate ? p[1] + acc[mi][ni][1] : acc[mi][ni][1];
    }
  }
}

int launch_gemm_tn_f64(cudaStream_t st, int m, int n, int k, const double* A, long lda,
                       const double* B, long ldb, double* C, long ldc, int accumulate, int batch,
                       long sA, long sB, long sC) {
  if (m <= 0 || n <= 0 || batch <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    RN_CHECK(cudaFuncSetAttribute(gemm_tn_f64_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    RN_CHECK(cudaFuncSetAttribute(gemm_tn_f64_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    attr_set = true;
  }
  const bool vec16 = (lda % 2 == 0) && (ldb % 2 == 0) && (sA % 2 == 0) && (sB % 2 == 0) &&
                     (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0);
  dim3 grid((unsigned)ceil_div(n, G_BN), (unsigned)ceil_div(m, G_BM), (unsigned)batch);
  const bool prof = gemm_profile_on();
  cudaEvent_t prof_begin = prof ? gemm_profile_begin(st) : nullptr;
  if (vec16)
    { RN_LAUNCH(gemm_tn_f64_kernel<true>, grid, G_THREADS, G_SMEM_BYTES, st, A, B, C, m, n, k, lda, ldb, ldc,
                                                                    sA, sB, sC, accumulate); rn::g_launches++; }
  else
    { RN_LAUNCH(gemm_tn_f64_kernel<false>, grid, G_THREADS, G_SMEM_BYTES, st, A, B, C, m, n, k, lda, ldb,
                                                                     ldc, sA, sB, sC, accumulate); rn::g_launches++; }
  if (prof_begin) gemm_profile_end(st, 0, prof_begin, 2.0 * m * n * k * batch);
  RN_LAUNCH_CHECK();
  return 0;
}

}  // namespace rn

extern "C" int rn_dgemm_tn(void* stream, int m, int n, int k, const double* A, long lda,
                           const double* B, long ldb, double* C, long ldc, int accumulate,
                           int batch, long strideA, long strideB, long strideC) {
  return rn::launch_gemm_tn_f64((cudaStream_t)stream, m, n, k, A, lda, B, ldb, C, ldc, accumulate,
                                batch, strideA, strideB, strideC);
}
