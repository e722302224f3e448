// Operand packing: strided (possibly transposed / conjugated / complex) tensor views -> K-major
// real operands for the GEMM kernels.
//
// The logical source matrix P has element (r, c) at src[r*s_row + c*s_col] (strides in ELEMENTS,
// an element being a double or an interleaved complex128).  c is the contraction (K) index.
//   A-form: dst[r*ld + c]                        real
//           dst[r*ld + 2c + {0,1}] = (re, +-im)  complex (conj flips the sign of im)
//   B-form (complex only, "realification" of the right operand so that a complex product becomes
//   one real GEMM with K -> 2K, N -> 2N):
//           dst[(2r  )*ld + 2c + {0,1}] = ( re, -im)      [conj_left: ( re, +im)]
//           dst[(2r+1)*ld + 2c + {0,1}] = ( im,  re)      [conj_left: ( im, -re)]
//   conj_left folds a complex conjugation of the LEFT operand into the right one.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

namespace rn {

template <bool CPLX>
struct Elt;
template <>
struct Elt<false> { using T = double; };
template <>
struct Elt<true> { using T = double2; };

template <bool CPLX, int MODE>
__global__ void __launch_bounds__(256)
pack_kernel(const typename Elt<CPLX>::T* __restrict__ src, double* __restrict__ dst, int rows,
            int cols, long s_row, long s_col, long dst_ld, int conj_flag, int c_fast) {
  pdl_wait();
  using T = typename Elt<CPLX>::T;
  __shared__ T tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int rl, cl;
    if (c_fast) { rl = ty + 8 * i; cl = tx; } else { cl = ty + 8 * i; rl = tx; }
    const int r = r0 + rl, c = c0 + cl;
    if (r < rows && c < cols) tile[rl][cl] = src[(long)r * s_row + (long)c * s_col];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 8 * i, cl = tx;
    const int r = r0 + rl, c = c0 + cl;
    if (r >= rows || c >= cols) continue;
    const T v = tile[rl][cl];
    if constexpr (!CPLX) {
      dst[(long)r * dst_ld + c] = v;
    } else if constexpr (MODE == 0) {
      double2 o = v;
      if (conj_flag) o.y = -o.y;
      *reinterpret_cast<double2*>(dst + (long)r * dst_ld + 2 * c) = o;
    } else {
      double2 o0, o1;
      if (!conj_flag) { o0 = make_double2(v.x, -v.y); o1 = make_double2(v.y, v.x); }
      else            { o0 = make_double2(v.x,  v.y); o1 = make_double2(v.y, -v.x); }
      *reinterpret_cast<double2*>(dst + (long)(2 * r) * dst_ld + 2 * c) = o0;
      *reinterpret_cast<double2*>(dst + (long)(2 * r + 1) * dst_ld + 2 * c) = o1;
    }
  }
}

int launch_pack(cudaStream_t st, int cplx, int mode, int conj_flag, int rows, int cols,
                const void* src, long s_row, long s_col, double* dst, long dst_ld) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!cplx && (mode != 0 || conj_flag)) return (int)cudaErrorInvalidValue;
  if (cplx && ((dst_ld & 1) || ((uintptr_t)dst & 15))) return (int)cudaErrorInvalidValue;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
  const int c_fast = s_col <= s_row;
  if (!cplx)
    { RN_LAUNCH((pack_kernel<false, 0>), grid, block, 0, st, (const double*)src, dst, rows, cols, s_row,
                                                   s_col, dst_ld, 0, c_fast); rn::g_launches++; }
  else if (mode == 0)
    { RN_LAUNCH((pack_kernel<true, 0>), grid, block, 0, st, (const double2*)src, dst, rows, cols, s_row,
                                                  s_col, dst_ld, conj_flag, c_fast); rn::g_launches++; }
  else
    { RN_LAUNCH((pack_kernel<true, 1>), grid, block, 0, st, (const double2*)src, dst, rows, cols, s_row,
                                                  s_col, dst_ld, conj_flag, c_fast); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

}  // namespace rn

extern "C" int rn_pack(void* stream, int cplx, int mode, int conj_flag, int rows, int cols,
                       const void* src, long s_row, long s_col, double* dst, long dst_ld) {
  return rn::launch_pack((cudaStream_t)stream, cplx, mode, conj_flag, rows, cols, src, s_row,
                         s_col, dst, dst_ld);
}
