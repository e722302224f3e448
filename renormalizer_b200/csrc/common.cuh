// Shared helpers for the renormalizer_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define RN_CHECK(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fprintf(stderr, "[rn_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),        \
              __FILE__, __LINE__, cudaGetErrorString(_e));                                   \
      return (int)_e;                                                                        \
    }                                                                                        \
  } while (0)

#define RN_LAUNCH_CHECK() RN_CHECK(cudaGetLastError())

namespace rn {

// ---- programmatic dependent launch -----------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerialization and starts with
// griddepcontrol.wait: the next kernel's CTAs are scheduled (and run their prologue) while the
// previous grid drains, instead of after the stream's full launch-to-launch gap.  The wait makes
// all memory of the prerequisite grids visible, so kernel bodies are unchanged.
// pdl_wait(): let the dependent grid be scheduled as soon as every CTA of this grid is resident
// (launch_dependents), then block until the prerequisite grids have completed and flushed.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() {
  pdl_trigger();
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static int pdl_on = -1;                 // RN_PDL=0 launches without the attribute (diagnostics)
  if (pdl_on < 0) { const char* e = getenv("RN_PDL"); pdl_on = (e && e[0] == '0') ? 0 : 1; }
  cfg.attrs = attr; cfg.numAttrs = pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define RN_LAUNCH(kernel, grid, block, smem, st, ...) \
  (void)rn::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, ##__VA_ARGS__)

extern long g_launches;   // kernels launched by this library (bench.py's gpu_launches)

__host__ __device__ inline long ceil_div(long a, long b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// cp.async with zero fill: copies src_bytes (<= CP bytes) and zero-fills the rest.
template <int CP>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "n"(CP), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV doubles per thread; result valid in every thread.  `scratch` must hold
// NV * 32 doubles.  Deterministic (fixed reduction tree).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = (lane < nwarp) ? scratch[i * 32 + lane] : 0.0;
    v[i] = warp_sum(t);
  }
}

}  // namespace rn
