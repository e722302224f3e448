// SVD of a bond matrix block, A = U diag(S) V^H, by one-sided (Hestenes) Jacobi iteration.
//   rn_svd        : QR-preconditioned (Drmac-Veselic) block Jacobi -- the product path (svd_qn)
//   rn_svd_jacobi : the bare iteration on the block itself (scalar kernel below 64 columns)
//
// Scalar kernel: columns are held as contiguous rows (At[c*ldt + r] = A[r][c]) so the three inner products and
// the plane rotation of a column pair are coalesced streams reduced with warp shuffles.  A sweep
// is n-1 rounds of a round-robin tournament; the n/2 disjoint pairs of one round run in parallel,
// one block per pair.  Rotations are accumulated into V the same way.  High relative accuracy of
// the small singular values (Demmel-Veselic) is why Jacobi is used for the truncation step.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"
#include <algorithm>
#include <vector>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace rn {

// cudaFreeAsync of every registered pointer when the scope is left, on success and on error alike
struct ScratchGuard {
  cudaStream_t st;
  std::vector<void**> ptrs;
  explicit ScratchGuard(cudaStream_t s) : st(s) {}
  void add(void** p) { ptrs.push_back(p); }
  ~ScratchGuard() { for (void** p : ptrs) if (*p) cudaFreeAsync(*p, st); }
};

constexpr int J_THREADS = 256;

template <bool CPLX>
__global__ void __launch_bounds__(J_THREADS)
jacobi_round_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                    typename std::conditional<CPLX, double2, double>::type* __restrict__ Vw,
                    int m, int n, int nv, long ldt, long ldv, int round, int N, double tol,
                    int* __restrict__ rotated) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ double scratch[4 * 32];
  int p, q;
  if (blockIdx.x == 0) { p = N - 1; q = round; }
  else { p = (round + blockIdx.x) % (N - 1); q = (round - (int)blockIdx.x + (N - 1)) % (N - 1); }
  if (p >= n || q >= n) return;
  if (p > q) { const int t = p; p = q; q = t; }
  T* xp = At + (long)p * ldt;
  T* xq = At + (long)q * ldt;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // |xp|^2, |xq|^2, Re<xp,xq>, Im<xp,xq>
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) {
      const double2 a = xp[r], b = xq[r];
      acc[0] += a.x * a.x + a.y * a.y;
      acc[1] += b.x * b.x + b.y * b.y;
      acc[2] += a.x * b.x + a.y * b.y;
      acc[3] += a.x * b.y - a.y * b.x;
    } else {
      const double a = xp[r], b = xq[r];
      acc[0] += a * a; acc[1] += b * b; acc[2] += a * b;
    }
  }
  block_sum<4>(acc, scratch);
  const double a = acc[0], b = acc[1];
  const double gabs = sqrt(acc[2] * acc[2] + acc[3] * acc[3]);
  if (gabs == 0.0 || gabs <= tol * sqrt(a) * sqrt(b)) return;
  if (threadIdx.x == 0) *rotated = 1;
  const double phr = acc[2] / gabs, phi = acc[3] / gabs;  // phase = g / |g|
  const double zeta = (b - a) / (2.0 * gabs);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  // xp' = c xp - s conj(phase) xq ;  xq' = s phase xp + c xq
  for (int pass = 0; pass < 2; ++pass) {
    T* up = pass == 0 ? xp : Vw + (long)p * ldv;
    T* uq = pass == 0 ? xq : Vw + (long)q * ldv;
    const int len = pass == 0 ? m : nv;
    for (int r = threadIdx.x; r < len; r += blockDim.x) {
      if constexpr (CPLX) {
        const double2 u = up[r], v = uq[r];
        // conj(phase) * v
        const double cvx = phr * v.x + phi * v.y, cvy = phr * v.y - phi * v.x;
        // phase * u
        const double pux = phr * u.x - phi * u.y, puy = phr * u.y + phi * u.x;
        up[r] = make_double2(c * u.x - s * cvx, c * u.y - s * cvy);
        uq[r] = make_double2(s * pux + c * v.x, s * puy + c * v.y);
      } else {
        const double u = up[r], v = uq[r];
        up[r] = c * u - s * phr * v;
        uq[r] = s * phr * u + c * v;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Block one-sided Jacobi on PAIRS OF COLUMN BLOCKS (2 x 16 columns), three kernels per round:
//
// The scalar kernel above needs n-1 launches per sweep, each touching every column once for three
// dot products and once for the rotation.  Here, for every block pair of a round,
//   1. jb_gram_kernel forms the Gram matrix G = Xc^H Xc of the 32 columns (row slices in parallel,
//      shared-memory row chunks, 4 entries of G per thread),
//   2. jb_rotate_kernel runs ONE cyclic sweep of two-sided Jacobi on G in shared memory (16 disjoint
//      pairs per round: 31 rounds in the first block round of a sweep, which also rotates the pairs
//      inside each block, 16 rounds joining the two blocks otherwise; thread (i, j) owns the 2x2
//      sub-block {p_i,q_i} x {p_j,q_j}, so a round is a rotation set-up by 16 threads and one
//      conflict-free update), accumulating J,
//   3. jb_apply_kernel applies Xc <- Xc J and Vc <- Vc J (row slices in parallel).
// In exact arithmetic this is the scalar algorithm with the pairs visited block by block (the
// rotation of a pair is computed from a, b, g exactly as above; G is updated by the same rotations
// instead of being recomputed from the columns).  Demmel & Veselic's relative accuracy of Jacobi on
// G = D A D carries over because G is formed from the current columns at every visit.  A sweep is
// 3 (n/16 - 1) launches instead of n - 1, and every column is read twice and written once per round.
constexpr int JB = 16;            // columns per block
constexpr int JK = 2 * JB;        // columns per CTA
constexpr int JBT = 256;          // threads

__device__ __forceinline__ double jb_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 jb_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double jb_cmul(double a, double b) { return a * b; }          // conj(a) * b
__device__ __forceinline__ double2 jb_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double jb_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 jb_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double jb_sub(double a, double b) { return a - b; }
__device__ __forceinline__ double2 jb_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double jb_scale(double c, double a) { return c * a; }
__device__ __forceinline__ double2 jb_scale(double c, double2 a) { return make_double2(c * a.x, c * a.y); }
__device__ __forceinline__ void jb_fma(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void jb_fma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void jb_cfma(double& acc, double a, double b) { acc = fma(a, b, acc); }   // += conj(a) b
__device__ __forceinline__ void jb_cfma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
template <typename T> __device__ __forceinline__ T jb_zero();
template <> __device__ __forceinline__ double jb_zero<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 jb_zero<double2>() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double jb_real(double a) { return a; }
__device__ __forceinline__ double jb_real(double2 a) { return a.x; }
__device__ __forceinline__ double jb_abs(double a) { return fabs(a); }
__device__ __forceinline__ double jb_abs(double2 a) { return sqrt(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ double jb_abs2(double a) { return a * a; }
__device__ __forceinline__ double jb_abs2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double jb_conj(double a) { return a; }
__device__ __forceinline__ double2 jb_conj(double2 a) { return make_double2(a.x, -a.y); }

template <bool CPLX> struct JBCfg { static constexpr int RC = CPLX ? 32 : 64; };   // rows per chunk
constexpr int JB_MAXSLICE = 8;    // row slices per block pair (CTAs working on the same 32 columns)

// block pair of CTA `bx` in round `round` of the round-robin tournament over NB (even) blocks
__device__ __forceinline__ void jb_pair(int bx, int round, int NB, int& P, int& Q) {
  if (bx == 0) { P = NB - 1; Q = round; }
  else { P = (round + bx) % (NB - 1); Q = (round - bx + (NB - 1)) % (NB - 1); }
  if (P > Q) { const int t = P; P = Q; Q = t; }
}
__device__ __forceinline__ int jb_col(int c, int P, int Q, int nb, int n) {
  const int blk = c < JB ? P : Q;
  const int idx = blk * JB + (c & (JB - 1));
  return (blk < nb && idx < n) ? idx : -1;
}

// 1. partial Gram matrices: CTA (pair, slice) -> Gp[(pair * nslice + slice) * JK * JK]
template <bool CPLX>
__global__ void __launch_bounds__(JBT)
jb_gram_kernel(const typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
               int m, int n, long ldt, int round, int NB, int nb, int rows_per_slice,
               typename std::conditional<CPLX, double2, double>::type* __restrict__ Gp) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int RC = JBCfg<CPLX>::RC, LDT = RC + 1;
  __shared__ T tile[JK * LDT];
  __shared__ int col[JK];
  const int tid = threadIdx.x;
  int P, Q;
  jb_pair(blockIdx.x, round, NB, P, Q);
  if (P >= nb) return;
  if (tid < JK) col[tid] = jb_col(tid, P, Q, nb, n);
  __syncthreads();
  const int rbeg = blockIdx.y * rows_per_slice;
  const int rend = min(m, rbeg + rows_per_slice);
  const int gi = tid >> 3, gj = (tid & 7) * 4;
  T acc[4] = {jb_zero<T>(), jb_zero<T>(), jb_zero<T>(), jb_zero<T>()};
  for (int r0 = rbeg; r0 < rend; r0 += RC) {
    for (int e = tid; e < JK * RC; e += JBT) {
      const int c = e / RC, r = e % RC, gr = r0 + r;
      const int cc = col[c];
      tile[c * LDT + r] = (cc >= 0 && gr < rend) ? At[(long)cc * ldt + gr] : jb_zero<T>();
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < RC; ++r) {
      const T xi = tile[gi * LDT + r];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) jb_cfma(acc[jj], xi, tile[(gj + jj) * LDT + r]);
    }
    __syncthreads();
  }
  T* out = Gp + ((long)blockIdx.x * gridDim.y + blockIdx.y) * (JK * JK);
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) out[gi * JK + gj + jj] = acc[jj];
}

// 2. G = sum of the slices (fixed order); one cyclic sweep of two-sided Jacobi on G in shared
//    memory; the accumulated rotation J and a "rotated" flag per pair go to global memory.
template <bool CPLX>
__global__ void __launch_bounds__(JBT)
jb_rotate_kernel(const typename std::conditional<CPLX, double2, double>::type* __restrict__ Gp, int nslice,
                 int round, int NB, int nb, double tol, int full,
                 typename std::conditional<CPLX, double2, double>::type* __restrict__ Jg,
                 int* __restrict__ pair_rot, int* __restrict__ rotated) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ T G[JK * JK];
  __shared__ T J[JK * JK];
  __shared__ int rp[JB], rq[JB];
  __shared__ double rc[JB];
  __shared__ T rs[JB];
  __shared__ int any_rot;
  const int tid = threadIdx.x;
  int P, Q;
  jb_pair(blockIdx.x, round, NB, P, Q);
  if (P >= nb) return;
  if (tid == 0) any_rot = 0;
  for (int e = tid; e < JK * JK; e += JBT) {
    T g = jb_zero<T>();
    for (int sl = 0; sl < nslice; ++sl) g = jb_add(g, Gp[((long)blockIdx.x * nslice + sl) * (JK * JK) + e]);
    G[e] = g;
    T one = jb_zero<T>();
    if ((e / JK) == (e % JK)) { if constexpr (CPLX) one.x = 1.0; else one = 1.0; }
    J[e] = one;
  }
  __syncthreads();
  const int ti = tid >> 4, tj = tid & 15;
  // full: all 496 pairs of the 32 columns (31 rounds).  Otherwise only the 256 pairs that join the
  // two blocks (16 rounds): the pairs inside a block are rotated once per sweep, in its first round,
  // where every block takes part in exactly one block pair.
  const int nrounds = full ? JK - 1 : JB;
  for (int lr = 0; lr < nrounds; ++lr) {
    if (tid < JB) {
      int p, q;
      if (!full) { p = tid; q = JB + ((tid + lr) & (JB - 1)); }
      else if (tid == 0) { p = JK - 1; q = lr; }
      else { p = (lr + tid) % (JK - 1); q = (lr - tid + (JK - 1)) % (JK - 1); }
      if (p > q) { const int t = p; p = q; q = t; }
      const double a = jb_real(G[p * JK + p]), b = jb_real(G[q * JK + q]);
      const T g = G[p * JK + q];
      const double g2 = jb_abs2(g);                   // |g|^2
      double c = 1.0;
      T sph = jb_zero<T>();
      // |g| > tol sqrt(a b), tested on the squares; the rotation with zeta = (b - a) / (2 |g|),
      // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) is evaluated as t = sign(d) 2|g| w with
      // w = 1 / (|d| + sqrt(d^2 + 4 |g|^2)), d = b - a: one square root, one reciprocal and one
      // reciprocal square root on the critical path of a round
      if (a > 0.0 && b > 0.0 && g2 > 0.0 && g2 > tol * tol * a * b) {
        const double d = b - a;
        const double w = 1.0 / (fabs(d) + sqrt(fma(d, d, 4.0 * g2)));
        const double tg = (d >= 0.0 ? 2.0 : -2.0) * w;     // t / |g|
        const double t2 = tg * tg * g2;                    // t^2
        c = rsqrt(1.0 + t2);
        sph = jb_scale(c * tg, g);                    // s * phase = c t g / |g|
        any_rot = 1;
      }
      rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = sph;
    }
    __syncthreads();
    {
      const int p = rp[ti], q = rq[ti], u = rp[tj], v = rq[tj];
      const double ci = rc[ti], cj = rc[tj];
      const T si = rs[ti], sj = rs[tj];
      const T gpu = G[p * JK + u], gpv = G[p * JK + v], gqu = G[q * JK + u], gqv = G[q * JK + v];
      // columns: xu' = cj xu - conj(sj) xv ; xv' = sj xu + cj xv
      const T pu = jb_sub(jb_scale(cj, gpu), jb_cmul(sj, gpv)), pv = jb_add(jb_mul(sj, gpu), jb_scale(cj, gpv));
      const T qu = jb_sub(jb_scale(cj, gqu), jb_cmul(sj, gqv)), qv = jb_add(jb_mul(sj, gqu), jb_scale(cj, gqv));
      // rows (R^H from the left): p' = ci p - si q ; q' = conj(si) p + ci q
      T npu = jb_sub(jb_scale(ci, pu), jb_mul(si, qu)), npv = jb_sub(jb_scale(ci, pv), jb_mul(si, qv));
      T nqu = jb_add(jb_cmul(si, pu), jb_scale(ci, qu)), nqv = jb_add(jb_cmul(si, pv), jb_scale(ci, qv));
      if (ti == tj) {                                  // the rotated pair itself: exactly diagonal
        npv = jb_zero<T>(); nqu = jb_zero<T>();
        if constexpr (CPLX) { npu.y = 0.0; nqv.y = 0.0; }
      }
      G[p * JK + u] = npu; G[p * JK + v] = npv; G[q * JK + u] = nqu; G[q * JK + v] = nqv;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = ti + h * JB;
        const T ju = J[rr * JK + u], jv = J[rr * JK + v];
        J[rr * JK + u] = jb_sub(jb_scale(cj, ju), jb_cmul(sj, jv));
        J[rr * JK + v] = jb_add(jb_mul(sj, ju), jb_scale(cj, jv));
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    pair_rot[blockIdx.x] = any_rot;
    if (any_rot) *rotated = 1;
  }
  if (!any_rot) return;
  T* out = Jg + (long)blockIdx.x * (JK * JK);
  for (int e = tid; e < JK * JK; e += JBT) out[e] = J[e];
}

// 3. Xc <- Xc J (blockIdx.z == 0) and Vc <- Vc J (blockIdx.z == 1), one CTA per row slice
template <bool CPLX>
__global__ void __launch_bounds__(JBT)
jb_apply_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                typename std::conditional<CPLX, double2, double>::type* __restrict__ Vw,
                int m, int n, int nv, long ldt, long ldv, int round, int NB, int nb,
                int rows_per_slice_x, int rows_per_slice_v,
                const typename std::conditional<CPLX, double2, double>::type* __restrict__ Jg,
                const int* __restrict__ pair_rot) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int RC = JBCfg<CPLX>::RC, LDT = RC + 1;
  __shared__ T tile[JK * LDT];
  __shared__ T J[JK * JK];
  __shared__ int col[JK];
  const int tid = threadIdx.x;
  int P, Q;
  jb_pair(blockIdx.x, round, NB, P, Q);
  if (P >= nb) return;
  if (!pair_rot[blockIdx.x]) return;
  if (tid < JK) col[tid] = jb_col(tid, P, Q, nb, n);
  for (int e = tid; e < JK * JK; e += JBT) J[e] = Jg[(long)blockIdx.x * (JK * JK) + e];
  __syncthreads();
  T* M = blockIdx.z == 0 ? At : Vw;
  const long ld = blockIdx.z == 0 ? ldt : ldv;
  const int rows = blockIdx.z == 0 ? m : nv;
  const int rps = blockIdx.z == 0 ? rows_per_slice_x : rows_per_slice_v;
  const int rbeg = blockIdx.y * rps;
  const int rend = min(rows, rbeg + rps);
  constexpr int NG = JBT / RC, NC = JK / NG;           // column groups, columns per thread
  const int r = tid % RC, cg = (tid / RC) * NC;
  for (int r0 = rbeg; r0 < rend; r0 += RC) {
    for (int e = tid; e < JK * RC; e += JBT) {
      const int c = e / RC, rr = e % RC, gr = r0 + rr;
      const int cc = col[c];
      tile[c * LDT + rr] = (cc >= 0 && gr < rend) ? M[(long)cc * ld + gr] : jb_zero<T>();
    }
    __syncthreads();
    T acc[NC];
#pragma unroll
    for (int jj = 0; jj < NC; ++jj) acc[jj] = jb_zero<T>();
#pragma unroll 4
    for (int c = 0; c < JK; ++c) {
      const T x = tile[c * LDT + r];
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) jb_fma(acc[jj], x, J[c * JK + cg + jj]);
    }
    if (r0 + r < rend) {
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int cc = col[cg + jj];
        if (cc >= 0) M[(long)cc * ld + r0 + r] = acc[jj];
      }
    }
    __syncthreads();
  }
}

// ---- one block round as ONE cluster launch -------------------------------------------------------
// The three kernels above exchange the partial Gram matrices and the rotation J through global
// memory and pay two launch boundaries per round; a round is ~35 us of which ~4 us are arithmetic.
// Here a cluster of S CTAs owns one block pair: CTA s keeps its row slice of the 32 columns of X and
// of V resident in shared memory, forms its partial Gram matrix, every CTA sums the S partials
// through DSMEM in a fixed order (so all CTAs hold the same G bit for bit), runs the same Jacobi
// sweep on it redundantly, and applies J to its own resident slices.  Columns are read once and
// written once per round, G and J never leave the SMs.
constexpr int JF_MAXROWS_BYTES = 160 * 1024;     // resident X + V slices per CTA

template <bool CPLX>
__global__ void __launch_bounds__(JBT)
jb_fused_round_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                      typename std::conditional<CPLX, double2, double>::type* __restrict__ Vw,
                      int m, int n, int nv, long ldt, long ldv, int round, int NB, int nb, int rps_x, int rps_v,
                      double tol, int full, int* __restrict__ rotated) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks();
  const int slice = (int)cluster.block_rank();
  const int pair = blockIdx.x / S;
  extern __shared__ __align__(16) unsigned char jf_smem[];
  T* Gp = reinterpret_cast<T*>(jf_smem);                 // my partial Gram matrix (read by the siblings)
  T* G = Gp + JK * JK;
  T* J = G + JK * JK;
  T* tx = J + JK * JK;                                   // X slice: [JK][rps_x + 1]
  const int ldx = rps_x + 1, ldvs = rps_v + 1;
  T* tv = tx + (long)JK * ldx;                           // V slice: [JK][rps_v + 1]
  __shared__ int col[JK];
  __shared__ int rp[JB], rq[JB];
  __shared__ double rc[JB];
  __shared__ T rs[JB];
  __shared__ int any_rot;
  const int tid = threadIdx.x;
  int P, Q;
  jb_pair(pair, round, NB, P, Q);
  const bool live = P < nb;                              // bye pair of an odd tournament: nothing to do
  if (tid < JK) col[tid] = live ? jb_col(tid, P, Q, nb, n) : -1;
  if (tid == 0) any_rot = 0;
  __syncthreads();
  const int xbeg = slice * rps_x, xend = min(m, xbeg + rps_x);
  const int vbeg = slice * rps_v, vend = min(nv, vbeg + rps_v);
  if (live) {
    for (int e = tid; e < JK * rps_x; e += JBT) {
      const int c = e / rps_x, r = e % rps_x, gr = xbeg + r;
      const int cc = col[c];
      tx[c * ldx + r] = (cc >= 0 && gr < xend) ? At[(long)cc * ldt + gr] : jb_zero<T>();
    }
    for (int e = tid; e < JK * rps_v; e += JBT) {
      const int c = e / rps_v, r = e % rps_v, gr = vbeg + r;
      const int cc = col[c];
      tv[c * ldvs + r] = (cc >= 0 && gr < vend) ? Vw[(long)cc * ldv + gr] : jb_zero<T>();
    }
  }
  __syncthreads();
  {
    // partial Gram matrix of my rows: thread -> row gi of G, 4 consecutive columns
    const int gi = tid >> 3, gj = (tid & 7) * 4;
    T acc[4] = {jb_zero<T>(), jb_zero<T>(), jb_zero<T>(), jb_zero<T>()};
    if (live) {
      const int rows = xend - xbeg;
      for (int r = 0; r < rows; ++r) {
        const T xi = tx[gi * ldx + r];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) jb_cfma(acc[jj], xi, tx[(gj + jj) * ldx + r]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) Gp[gi * JK + gj + jj] = acc[jj];
  }
  cluster.sync();
  // G = sum of the partials in slice order (identical in every CTA) ; J = I
  for (int e = tid; e < JK * JK; e += JBT) {
    T g = jb_zero<T>();
    for (int sl = 0; sl < S; ++sl) {
      const T* remote = cluster.map_shared_rank(Gp, sl);
      g = jb_add(g, remote[e]);
    }
    G[e] = g;
    T one = jb_zero<T>();
    if ((e / JK) == (e % JK)) { if constexpr (CPLX) one.x = 1.0; else one = 1.0; }
    J[e] = one;
  }
  __syncthreads();
  // pairs this visit would rotate (all 496, or the 256 that join the two blocks): when none of them
  // exceeds the threshold -- most block pairs of the last sweeps -- the rotation rounds are skipped
  int need = 0;
  if (live) {
    for (int e = tid; e < JK * JK; e += JBT) {
      const int i = e / JK, j = e % JK;
      if (i < j && (full || (i < JB && j >= JB))) {
        const double a = jb_real(G[i * JK + i]), b = jb_real(G[j * JK + j]);
        const double g2 = jb_abs2(G[e]);
        if (a > 0.0 && b > 0.0 && g2 > 0.0 && g2 > tol * tol * a * b) need = 1;
      }
    }
  }
  need = __syncthreads_or(need);
  if (live && need) {
    const int ti = tid >> 4, tj = tid & 15;
    const int nrounds = full ? JK - 1 : JB;
    for (int lr = 0; lr < nrounds; ++lr) {
      if (tid < JB) {
        int p, q;
        if (!full) { p = tid; q = JB + ((tid + lr) & (JB - 1)); }
        else if (tid == 0) { p = JK - 1; q = lr; }
        else { p = (lr + tid) % (JK - 1); q = (lr - tid + (JK - 1)) % (JK - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        const double a = jb_real(G[p * JK + p]), b = jb_real(G[q * JK + q]);
        const T g = G[p * JK + q];
        const double g2 = jb_abs2(g);
        double c = 1.0;
        T sph = jb_zero<T>();
        if (a > 0.0 && b > 0.0 && g2 > 0.0 && g2 > tol * tol * a * b) {
          const double d = b - a;
          const double w = 1.0 / (fabs(d) + sqrt(fma(d, d, 4.0 * g2)));
          const double tg = (d >= 0.0 ? 2.0 : -2.0) * w;
          const double t2 = tg * tg * g2;
          c = rsqrt(1.0 + t2);
          sph = jb_scale(c * tg, g);
          any_rot = 1;
        }
        rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = sph;
      }
      __syncthreads();
      {
        const int p = rp[ti], q = rq[ti], u = rp[tj], v = rq[tj];
        const double ci = rc[ti], cj = rc[tj];
        const T si = rs[ti], sj = rs[tj];
        const T gpu = G[p * JK + u], gpv = G[p * JK + v], gqu = G[q * JK + u], gqv = G[q * JK + v];
        const T pu = jb_sub(jb_scale(cj, gpu), jb_cmul(sj, gpv)), pv = jb_add(jb_mul(sj, gpu), jb_scale(cj, gpv));
        const T qu = jb_sub(jb_scale(cj, gqu), jb_cmul(sj, gqv)), qv = jb_add(jb_mul(sj, gqu), jb_scale(cj, gqv));
        T npu = jb_sub(jb_scale(ci, pu), jb_mul(si, qu)), npv = jb_sub(jb_scale(ci, pv), jb_mul(si, qv));
        T nqu = jb_add(jb_cmul(si, pu), jb_scale(ci, qu)), nqv = jb_add(jb_cmul(si, pv), jb_scale(ci, qv));
        if (ti == tj) {
          npv = jb_zero<T>(); nqu = jb_zero<T>();
          if constexpr (CPLX) { npu.y = 0.0; nqv.y = 0.0; }
        }
        G[p * JK + u] = npu; G[p * JK + v] = npv; G[q * JK + u] = nqu; G[q * JK + v] = nqv;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = ti + h * JB;
          const T ju = J[rr * JK + u], jv = J[rr * JK + v];
          J[rr * JK + u] = jb_sub(jb_scale(cj, ju), jb_cmul(sj, jv));
          J[rr * JK + v] = jb_add(jb_mul(sj, ju), jb_scale(cj, jv));
        }
      }
      __syncthreads();
    }
    if (any_rot) {
      if (tid == 0 && slice == 0) *rotated = 1;
      // my slices <- slices . J : thread -> row r (of X, then of V), NC columns at a time from registers
      for (int which = 0; which < 2; ++which) {
        T* tile = which == 0 ? tx : tv;
        const int ld = which == 0 ? ldx : ldvs;
        const int rows = which == 0 ? xend - xbeg : vend - vbeg;
        T* M = which == 0 ? At : Vw;
        const long ldg = which == 0 ? ldt : ldv;
        const int rbeg = which == 0 ? xbeg : vbeg;
        // item = (row r, group of 8 output columns): consecutive threads -> consecutive rows
        for (int item = tid; item < rows * (JK / 8); item += JBT) {
          const int r = item % rows, j0 = (item / rows) * 8;
          T acc[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[jj] = jb_zero<T>();
#pragma unroll 4
          for (int c = 0; c < JK; ++c) {
            const T x = tile[c * ld + r];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) jb_fma(acc[jj], x, J[c * JK + j0 + jj]);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int cc = col[j0 + jj];
            if (cc >= 0) M[(long)cc * ldg + rbeg + r] = acc[jj];
          }
        }
      }
    }
  }
  cluster.sync();          // no CTA may exit while a sibling can still read its partial Gram matrix
}

// S[c] = |At[c]|, At[c] /= S[c]
template <bool CPLX>
__global__ void __launch_bounds__(J_THREADS)
jacobi_finalize_kernel(typename std::conditional<CPLX, double2, double>::type* __restrict__ At,
                       int m, long ldt, double* __restrict__ S) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ double scratch[32];
  T* x = At + (long)blockIdx.x * ldt;
  double ss[1] = {0.0};
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) { const double2 a = x[r]; ss[0] += a.x * a.x + a.y * a.y; }
    else { const double a = x[r]; ss[0] += a * a; }
  }
  block_sum<1>(ss, scratch);
  const double sig = sqrt(ss[0]);
  if (threadIdx.x == 0) S[blockIdx.x] = sig;
  const double inv = sig > 0.0 ? 1.0 / sig : 0.0;
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) { double2 a = x[r]; a.x *= inv; a.y *= inv; x[r] = a; }
    else x[r] *= inv;
  }
}

template <bool CPLX>
__global__ void jacobi_eye_kernel(typename std::conditional<CPLX, double2, double>::type* Vw, int n, long ldv) {
  pdl_wait();
  const long total = (long)n * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / n), r = (int)(i % n);
    if constexpr (CPLX) Vw[(long)c * ldv + r] = make_double2(r == c ? 1.0 : 0.0, 0.0);
    else Vw[(long)c * ldv + r] = r == c ? 1.0 : 0.0;
  }
}

// Jacobi iteration on nt columns of length mt held as rows of At (ldt); on return row c of At is the
// normalised left vector u_c, S[c] its singular value and row c of Vw (nt x nt, ldv) column c of V.
template <bool CPLX>
static int jacobi_core(cudaStream_t st, int mt, int nt, typename std::conditional<CPLX, double2, double>::type* At,
                       long ldt, typename std::conditional<CPLX, double2, double>::type* Vw, long ldv,
                       double* S, int max_sweeps, int* sweeps_out) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  int* flag = nullptr;
  ScratchGuard guard(st);
  guard.add((void**)&flag);
  RN_CHECK(cudaMallocAsync((void**)&flag, sizeof(int), st));
  int nbe = (int)ceil_div((long)nt * nt, 256);
  if (nbe > 1184) nbe = 1184;
  { RN_LAUNCH(jacobi_eye_kernel<CPLX>, nbe, 256, 0, st, Vw, nt, ldv); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  const int N = (nt + 1) & ~1;  // even number of players
  const double tol = sqrt((double)mt) * 2.220446049250313e-16;
  int sweeps = 0;
  bool converged = nt <= 1;
  static int block_on = -1;                 // RN_SVD_BLOCK=0 keeps the scalar kernel (diagnostics)
  if (block_on < 0) {
    const char* e = getenv("RN_SVD_BLOCK");
    block_on = (e && e[0] == '0') ? 0 : 1;
  }
  const bool blocked = block_on && nt >= 2 * JK;
  const int nb = (int)ceil_div(nt, JB), NB = (nb + 1) & ~1;
  constexpr int RC = JBCfg<CPLX>::RC;
  // row slices: whole chunks of RC rows, at most JB_MAXSLICE CTAs per block pair
  auto slices = [&](int rows, int& rps) {
    int ns = (int)ceil_div(rows, RC);
    if (ns > JB_MAXSLICE) ns = JB_MAXSLICE;
    rps = (int)ceil_div(ceil_div(rows, ns), RC) * RC;
    return (int)ceil_div(rows, rps);
  };
  int rps_x = 0, rps_v = 0;
  const int ns_x = slices(mt, rps_x), ns_v = slices(nt, rps_v);
  // fused cluster round: S CTAs per block pair, X and V row slices resident in shared memory
  static int fused_on = -1;                 // RN_SVD_FUSED=0 keeps the three-kernel round (diagnostics)
  if (fused_on < 0) { const char* e = getenv("RN_SVD_FUSED"); fused_on = (e && e[0] == '0') ? 0 : 1; }
  int fS = 8;
  while (fS > 1 && ((mt + fS - 1) / fS < 16)) fS >>= 1;
  const int frps_x = (int)ceil_div(mt, fS), frps_v = (int)ceil_div(nt, fS);
  const size_t fsmem = sizeof(T) * ((size_t)3 * JK * JK + (size_t)JK * (frps_x + 1) + (size_t)JK * (frps_v + 1));
  // worth it only while every cluster of a round is resident at once (large blocks need most of an
  // SM's shared memory per CTA and would run in several waves: 2048 x 2048 measured 403 vs 234 ms)
  const long fused_ctas = (long)(NB / 2) * fS;
  const long per_sm = (long)(227 * 1024) / (long)(fsmem + 1024);
  bool fused_ok = blocked && fused_on && fsmem <= (size_t)JF_MAXROWS_BYTES + sizeof(T) * 3 * JK * JK &&
                  per_sm >= 1 && fused_ctas <= 148 * per_sm;
  if (fused_ok) {
    static size_t attr_bytes[2] = {0, 0};
    if (fsmem > 48 * 1024 && fsmem > attr_bytes[CPLX ? 1 : 0]) {
      RN_CHECK(cudaFuncSetAttribute(jb_fused_round_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(JF_MAXROWS_BYTES + sizeof(T) * 3 * JK * JK)));
      attr_bytes[CPLX ? 1 : 0] = JF_MAXROWS_BYTES + sizeof(T) * 3 * JK * JK;
    }
  }
  T *Gp = nullptr, *Jg = nullptr;
  int* pair_rot = nullptr;
  guard.add((void**)&Gp); guard.add((void**)&Jg); guard.add((void**)&pair_rot);
  if (blocked && !fused_ok) {
    RN_CHECK(cudaMallocAsync((void**)&Gp, sizeof(T) * (size_t)(NB / 2) * ns_x * JK * JK, st));
    RN_CHECK(cudaMallocAsync((void**)&Jg, sizeof(T) * (size_t)(NB / 2) * JK * JK, st));
    RN_CHECK(cudaMallocAsync((void**)&pair_rot, sizeof(int) * (size_t)(NB / 2), st));
  }
  if (nt > 1) {
    for (; sweeps < max_sweeps; ++sweeps) {
      RN_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
      if (blocked && fused_ok) {
        for (int round = 0; round < NB - 1; ++round) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3((unsigned)((NB / 2) * fS)); cfg.blockDim = dim3(JBT);
          cfg.dynamicSmemBytes = fsmem; cfg.stream = st;
          cudaLaunchAttribute attr[2];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = (unsigned)fS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[1].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = attr; cfg.numAttrs = 2;
          RN_CHECK(cudaLaunchKernelEx(&cfg, jb_fused_round_kernel<CPLX>, At, Vw, mt, nt, nt, ldt, ldv, round, NB, nb,
                                      frps_x, frps_v, tol, round == 0 ? 1 : 0, flag));
          rn::g_launches++;
        }
      } else if (blocked) {
        for (int round = 0; round < NB - 1; ++round) {
          RN_LAUNCH(jb_gram_kernel<CPLX>, dim3(NB / 2, ns_x), JBT, 0, st, At, mt, nt, ldt, round, NB, nb, rps_x, Gp);
          RN_LAUNCH(jb_rotate_kernel<CPLX>, NB / 2, JBT, 0, st, Gp, ns_x, round, NB, nb, tol, round == 0 ? 1 : 0, Jg,
                    pair_rot, flag);
          RN_LAUNCH(jb_apply_kernel<CPLX>, dim3(NB / 2, ns_x > ns_v ? ns_x : ns_v, 2), JBT, 0, st, At, Vw, mt, nt, nt,
                    ldt, ldv, round, NB, nb, rps_x, rps_v, Jg, pair_rot);
          rn::g_launches += 3;
        }
      } else {
        for (int round = 0; round < N - 1; ++round)
          { RN_LAUNCH(jacobi_round_kernel<CPLX>, N / 2, J_THREADS, 0, st, At, Vw, mt, nt, nt, ldt, ldv, round, N, tol, flag); rn::g_launches++; }
      }
      RN_LAUNCH_CHECK();
      int h = 0;
      RN_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      RN_CHECK(cudaStreamSynchronize(st));
      if (!h) { ++sweeps; converged = true; break; }
    }
  }
  // a negative count reports that the last sweep still rotated (not converged in max_sweeps)
  if (sweeps_out) *sweeps_out = converged ? sweeps : -sweeps;
  { RN_LAUNCH(jacobi_finalize_kernel<CPLX>, nt, J_THREADS, 0, st, At, mt, ldt, S); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  return 0;
}

// Economic SVD of A (m x n row-major, lda): U (m x k, ldu), S (k), Vh (k x n, ldvh), k = min(m,n).
// Singular values are NOT sorted.  Returns the number of sweeps in *sweeps_out (host int).
template <bool CPLX>
static int svd_driver(cudaStream_t st, int m, int n, const void* A, long lda, void* U, long ldu,
                      double* S, void* Vh, long ldvh, int max_sweeps, int* sweeps_out) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const int es = CPLX ? 2 : 1;
  const bool wide = m < n;
  const int mt = wide ? n : m;   // tall problem: mt x nt
  const int nt = wide ? m : n;
  const long ldt = mt, ldv = nt;
  T *At = nullptr, *Vw = nullptr;
  ScratchGuard guard(st);
  guard.add((void**)&At); guard.add((void**)&Vw);
  RN_CHECK(cudaMallocAsync((void**)&At, sizeof(T) * (size_t)nt * ldt, st));
  RN_CHECK(cudaMallocAsync((void**)&Vw, sizeof(T) * (size_t)nt * ldv, st));
  int err;
  if (!wide) err = launch_pack(st, CPLX, 0, 0, n, m, A, 1, lda, (double*)At, ldt * es);   // At = A^T
  else err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, m, n, A, lda, 1, (double*)At, ldt * es);  // At = conj(A)
  if (err) return err;
  if ((err = jacobi_core<CPLX>(st, mt, nt, At, ldt, Vw, ldv, S, max_sweeps, sweeps_out))) return err;
  if (!wide) {
    // U[r][c] = At[c][r];  Vh[c][j] = conj(V[j][c]) = conj(Vw[c][j])
    err = launch_pack(st, CPLX, 0, 0, m, nt, At, 1, ldt, (double*)U, ldu * es);
    if (err) return err;
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, nt, n, Vw, ldv, 1, (double*)Vh, ldvh * es);
  } else {
    // A = V' S U'^H:  U[j][c] = Vw[c][j];  Vh[c][r] = conj(U'[r][c]) = conj(At[c][r])
    err = launch_pack(st, CPLX, 0, 0, m, nt, Vw, 1, ldv, (double*)U, ldu * es);
    if (err) return err;
    err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, nt, n, At, ldt, 1, (double*)Vh, ldvh * es);
  }
  return err;
}

// ---------------------------------------------------------------------------------------------
// Preconditioned SVD (Drmac & Veselic, SIAM J. Matrix Anal. Appl. 29, 1322): with T the tall
// orientation of A (T = A or A^H, mt x k) and P the ordering of T's columns by decreasing norm,
//     T P = Q1 R1,   R1^H = Q2 R2,   X = R2^H = Ux S Vx^H  (Jacobi on the k x k factor)
//     =>  T P = (Q1 Ux) S (Q2 Vx)^H.
// Bond matrices of a converged state are strongly graded; the bare iteration on T does not converge
// in 40 sweeps on them and loses the orthogonality of the vectors of the small singular values,
// while X needs 6-9 sweeps over columns of length k.
template <bool CPLX>
__global__ void __launch_bounds__(J_THREADS)
col_norm2_kernel(const typename std::conditional<CPLX, double2, double>::type* __restrict__ At, int m, long ldt,
                 double* __restrict__ out) {
  pdl_wait();
  __shared__ double scratch[32];
  const auto* x = At + (long)blockIdx.x * ldt;
  double ss[1] = {0.0};
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    if constexpr (CPLX) { const double2 a = x[r]; ss[0] += a.x * a.x + a.y * a.y; }
    else { const double a = x[r]; ss[0] += a * a; }
  }
  block_sum<1>(ss, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = ss[0];
}

// dst[r * ldd + j] = src[perm[j] * lds + r]   (rows of `src` gathered and transposed), 32 x 32 tiles
template <bool CPLX>
__global__ void __launch_bounds__(256)
gather_rows_t_kernel(const typename std::conditional<CPLX, double2, double>::type* __restrict__ src, long lds,
                     const int* __restrict__ perm, int nrows_src, int len,
                     typename std::conditional<CPLX, double2, double>::type* __restrict__ dst, long ldd) {
  pdl_wait();
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j0 = blockIdx.y * 32, r0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty + 8 * i, r = r0 + tx;
    if (j < nrows_src && r < len) tile[ty + 8 * i][tx] = src[(long)perm[j] * lds + r];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, j = j0 + tx;
    if (j < nrows_src && r < len) dst[(long)r * ldd + j] = tile[tx][ty + 8 * i];
  }
}

// mode 0: dst[c * ldd + perm[j]] = conj(src[j * k + c])   (Vh of a tall problem)
// mode 1: dst[perm[j] * ldd + c] = src[j * k + c]         (U of a wide problem)
template <bool CPLX>
__global__ void __launch_bounds__(256)
perm_scatter_kernel(const typename std::conditional<CPLX, double2, double>::type* __restrict__ src, int k,
                    const int* __restrict__ perm, int mode,
                    typename std::conditional<CPLX, double2, double>::type* __restrict__ dst, long ldd) {
  pdl_wait();
  const long total = (long)k * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int j, c;
    if (mode == 0) { c = (int)(i / k); j = (int)(i % k); }     // consecutive threads: consecutive j
    else { j = (int)(i / k); c = (int)(i % k); }
    auto v = src[(long)j * k + c];
    if (mode == 0) {
      if constexpr (CPLX) v.y = -v.y;
      dst[(long)c * ldd + perm[j]] = v;
    } else {
      dst[(long)perm[j] * ldd + c] = v;
    }
  }
}

template <bool CPLX>
static int svd_precond_driver(cudaStream_t st, int m, int n, const void* A, long lda, void* U, long ldu,
                              double* S, void* Vh, long ldvh, int max_sweeps, int path, int* sweeps_out) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const int es = CPLX ? 2 : 1;
  const bool wide = m < n;
  const int mt = wide ? n : m, k = wide ? m : n;
  T *Tt = nullptr, *Ts = nullptr, *Q1 = nullptr, *R1 = nullptr, *B = nullptr, *Q2 = nullptr, *R2 = nullptr,
    *Vw = nullptr, *W = nullptr, *UT = nullptr;
  double* nrm = nullptr;
  int* perm = nullptr;
  ScratchGuard guard(st);
  for (void** q : {(void**)&Tt, (void**)&Ts, (void**)&Q1, (void**)&UT, (void**)&R1, (void**)&B, (void**)&Q2, (void**)&R2,
                   (void**)&Vw, (void**)&W, (void**)&nrm, (void**)&perm})
    guard.add(q);
  const size_t tall = sizeof(T) * (size_t)mt * k, sq = sizeof(T) * (size_t)k * k;
  RN_CHECK(cudaMallocAsync((void**)&Tt, tall, st));
  RN_CHECK(cudaMallocAsync((void**)&Ts, tall, st));
  RN_CHECK(cudaMallocAsync((void**)&Q1, tall, st));
  RN_CHECK(cudaMallocAsync((void**)&UT, tall, st));
  RN_CHECK(cudaMallocAsync((void**)&R1, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&B, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&Q2, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&R2, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&Vw, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&W, sq, st));
  RN_CHECK(cudaMallocAsync((void**)&nrm, sizeof(double) * k, st));
  RN_CHECK(cudaMallocAsync((void**)&perm, sizeof(int) * k, st));
  int err;
  // Tt[c][r] = T[r][c]: rows of Tt are the columns of the tall orientation
  if (!wide) err = launch_pack(st, CPLX, 0, 0, n, m, A, 1, lda, (double*)Tt, (long)mt * es);
  else err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, m, n, A, lda, 1, (double*)Tt, (long)mt * es);
  if (err) return err;
  { RN_LAUNCH(col_norm2_kernel<CPLX>, k, J_THREADS, 0, st, Tt, mt, (long)mt, nrm); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  {
    std::vector<double> hn(k);
    std::vector<int> hp(k);
    RN_CHECK(cudaMemcpyAsync(hn.data(), nrm, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
    RN_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < k; ++i) hp[i] = i;
    std::stable_sort(hp.begin(), hp.end(), [&](int a, int b) { return hn[a] > hn[b]; });
    RN_CHECK(cudaMemcpyAsync(perm, hp.data(), sizeof(int) * k, cudaMemcpyHostToDevice, st));
    RN_CHECK(cudaStreamSynchronize(st));          // hp goes out of scope
  }
  // Ts = T P (row-major mt x k)
  { RN_LAUNCH(gather_rows_t_kernel<CPLX>, dim3((unsigned)ceil_div(mt, 32), (unsigned)ceil_div(k, 32)), 256, 0, st,
              Tt, (long)mt, perm, k, mt, Ts, (long)k); rn::g_launches++; }
  RN_LAUNCH_CHECK();
  if ((err = rn_qr(st, CPLX, mt, k, Ts, k, Q1, k, R1, k))) return err;                       // T P = Q1 R1
  static int qr2_on = -1;                   // RN_SVD_QR2=0: one factorisation only, Jacobi on R1^H (diagnostics)
  if (qr2_on < 0) { const char* e = getenv("RN_SVD_QR2"); qr2_on = (e && e[0] == '0') ? 0 : 1; }
  T* ut_out = (!wide && ldu == k) ? (T*)U : UT;
  T* VTs = nullptr;                         // right vectors of T P, row-major k x k
  if (qr2_on) {
    if ((err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, k, k, R1, 1, k, (double*)B, (long)k * es))) return err;   // B = R1^H
    if ((err = rn_qr(st, CPLX, k, k, B, k, Q2, k, R2, k))) return err;                       // R1^H = Q2 R2
    // columns of X = R2^H are the conjugated rows of R2: Xt = conj(R2), same layout
    T* Xt = R2;
    if (CPLX) { if ((err = launch_pack(st, 1, 0, 1, k, k, R2, k, 1, (double*)B, (long)k * es))) return err; Xt = B; }
    if ((err = jacobi_core<CPLX>(st, k, k, Xt, (long)k, Vw, (long)k, S, max_sweeps, sweeps_out))) return err;
    // U_T = Q1 Ux, Ux[r][c] = Xt[c][r];   V_Ts = Q2 Vx, Vx[r][c] = Vw[c][r]
    if ((err = launch_pack(st, CPLX, 0, 0, k, k, Xt, 1, k, (double*)W, (long)k * es))) return err;
    if ((err = rn_matmul(st, CPLX, mt, k, k, Q1, W, ut_out, path))) return err;
    if ((err = launch_pack(st, CPLX, 0, 0, k, k, Vw, 1, k, (double*)W, (long)k * es))) return err;
    if ((err = rn_matmul(st, CPLX, k, k, k, Q2, W, R1, path))) return err;                    // R1 <- V_Ts
    VTs = R1;
  } else {
    // X = R1^H = Ux S Vx^H  =>  T P = (Q1 Vx) S Ux^H;  Xt = conj(R1), same layout
    T* Xt = R1;
    if (CPLX) { if ((err = launch_pack(st, 1, 0, 1, k, k, R1, k, 1, (double*)B, (long)k * es))) return err; Xt = B; }
    if ((err = jacobi_core<CPLX>(st, k, k, Xt, (long)k, Vw, (long)k, S, max_sweeps, sweeps_out))) return err;
    if ((err = launch_pack(st, CPLX, 0, 0, k, k, Vw, 1, k, (double*)W, (long)k * es))) return err;
    if ((err = rn_matmul(st, CPLX, mt, k, k, Q1, W, ut_out, path))) return err;               // U_T = Q1 Vx
    if ((err = launch_pack(st, CPLX, 0, 0, k, k, Xt, 1, k, (double*)R2, (long)k * es))) return err;   // V_Ts = Ux
    VTs = R2;
  }
  int nbs = (int)ceil_div((long)k * k, 256);
  if (nbs > 1184) nbs = 1184;
  if (!wide) {
    if (ut_out != (T*)U && (err = launch_pack(st, CPLX, 0, 0, m, k, UT, k, 1, (double*)U, ldu * es))) return err;
    { RN_LAUNCH(perm_scatter_kernel<CPLX>, nbs, 256, 0, st, VTs, k, perm, 0, (T*)Vh, ldvh); rn::g_launches++; }
  } else {
    // A = T^H = V_T S U_T^H:  U[perm[j]][c] = V_Ts[j][c];  Vh[c][r] = conj(U_T[r][c])
    { RN_LAUNCH(perm_scatter_kernel<CPLX>, nbs, 256, 0, st, VTs, k, perm, 1, (T*)U, ldu); rn::g_launches++; }
    if ((err = launch_pack(st, CPLX, 0, CPLX ? 1 : 0, k, mt, UT, 1, k, (double*)Vh, ldvh * es))) return err;
  }
  RN_LAUNCH_CHECK();
  return 0;
}

}  // namespace rn

extern "C" int rn_svd_jacobi(void* stream, int cplx, int m, int n, const void* A, long lda,
                             void* U, long ldu, double* S, void* Vh, long ldvh, int max_sweeps,
                             int* sweeps_out) {
  if (m <= 0 || n <= 0) return 0;
  if (max_sweeps <= 0) max_sweeps = 40;
  return cplx ? rn::svd_driver<true>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                     max_sweeps, sweeps_out)
              : rn::svd_driver<false>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                      max_sweeps, sweeps_out);
}

extern "C" int rn_svd(void* stream, int cplx, int m, int n, const void* A, long lda, void* U, long ldu,
                      double* S, void* Vh, long ldvh, int max_sweeps, int path, int* sweeps_out) {
  if (m <= 0 || n <= 0) return 0;
  if (max_sweeps <= 0) max_sweeps = 40;
  return cplx ? rn::svd_precond_driver<true>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                             max_sweeps, path, sweeps_out)
              : rn::svd_precond_driver<false>((cudaStream_t)stream, m, n, A, lda, U, ldu, S, Vh, ldvh,
                                              max_sweeps, path, sweeps_out);
}
