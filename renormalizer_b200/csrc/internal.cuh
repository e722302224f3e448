// Internal launcher declarations shared between the translation units of librn_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace rn {

struct WApplyParams {
  const void* in;
  void* out;
  int X, P, Q, Y;
  long isx, isp, isq, isy;
  int D, F, Y2;
  long osx, osd, osf, osy1, osy2;
  const int* rowptr;
  const int* ent_pq;
  const double* ent_val;
  int YT;
  int order;  // innermost-in-memory input index: 0 = y, 1 = p (then y), 2 = q (then y)
};

int launch_gemm_tn_f64(cudaStream_t st, int m, int n, int k, const double* A, long lda,
                       const double* B, long ldb, double* C, long ldc, int accumulate, int batch,
                       long sA, long sB, long sC);
int launch_pack(cudaStream_t st, int cplx, int mode, int conj_flag, int rows, int cols,
                const void* src, long s_row, long s_col, double* dst, long dst_ld);
int launch_wapply(cudaStream_t st, int cplx, const WApplyParams& p);

size_t ozaki_split_bytes(int rows, int K, int nslices);
int launch_ozaki_split(cudaStream_t st, const double* X, long ld, int rows, int K, int nslices,
                       signed char* q, double* scale);
int launch_ozaki_split_bform(cudaStream_t st, const double* X, long ld, int crows, int ccols, int nslices,
                             int conj_left, signed char* q, double* scale);
int launch_ozaki_split_t(cudaStream_t st, int cplx, const void* src, long s_col, int rows, int cols,
                         int nslices, signed char* q, double* scale);
int launch_wapply_split(cudaStream_t st, int cplx, const WApplyParams& p, int nslices, signed char* q,
                        double* scale);
int ozaki_make_map(CUtensorMap* map, const signed char* q, long total_rows, int Kp, int box_rows = 128);
int launch_ozaki_gemm_maps(cudaStream_t st, int m, int n, int K, int nslices, const CUtensorMap* tmA,
                           const double* sA, const CUtensorMap* tmB, const CUtensorMap* tmB64, const double* sB,
                           double* C, long ldc, const double* dotv = nullptr, double* dot_partial = nullptr);
int ozaki_gemm_tiles(int m, int n);
bool gemm_profile_on();
cudaEvent_t gemm_profile_begin(cudaStream_t st);
void gemm_profile_end(cudaStream_t st, int kind, cudaEvent_t begin, double flops);
int launch_ozaki_gemm(cudaStream_t st, int m, int n, int K, int nslices, const signed char* qA,
                      const double* sA, const signed char* qB, const double* sB, double* C, long ldc);

}  // namespace rn

struct rn_hop_plan;
namespace rn {
int hop_dot_tiles(const rn_hop_plan* p);
int hop_apply_dot(rn_hop_plan* p, cudaStream_t st, const void* c_in, void* out, double* dot_partial);
}  // namespace rn
