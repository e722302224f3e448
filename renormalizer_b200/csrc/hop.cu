// H_eff * C plans and the environment update: the two contraction hot paths of a sweep site,
// lowered onto {pack, GEMM, MPO-apply, GEMM}.
//
//   hop (reference renormalizer/mps/hop_expr.py:7-117)
//     0 site : out[a,l]        = L[a,b,c] R[l,b,k] C[c,k]
//     1 site : out[a,d,(g),l]  = L[a,b,c] W[b,d,e,f] R[l,f,k] C[c,e,(g),k]
//     2 sites: out[a,d,(m),g,(n),l] = L[a,b,c] W1[b,d,e,f] W2[f,g,h,j] R[l,j,k] C[c,e,(m),h,(n),k]
//   as   G1: T1[(a,b),(rest)] = L[(a,b),c] . C[c,(rest)]          (GEMM, K = |c|)
//        W : T2 = W1 applied on (b,e)  [, T3 = W2 applied on (f,h)] (HBM-bound, wapply.cu)
//        G3: out[(a,d..),l]   = T[(a,d..),(f,k)] . R[l,(f,k)]      (GEMM, K = |f||k|)
//   env update (reference renormalizer/mps/lib.py:172-262, contract_one_site) reuses G1 + W and
//   finishes with the bra-site GEMM.
//
// Complex tensors are contracted with REAL GEMMs: the left operand is used through its
// interleaved real view and the right operand is packed once into its 2x2 real representation
// (pack.cu, B-form).  L and R are constant over all Krylov / Davidson iterations of a site, so
// their packed forms live in the plan.
#include "common.cuh"
#include "rn_b200.h"
#include "internal.cuh"

#include <new>
#include <vector>

namespace rn {

struct WCsr {
  int P = 0, Q = 0, D = 0, F = 0;
  const int* rowptr = nullptr;
  const int* pq = nullptr;
  const double* val = nullptr;
};

// The int32 accumulators of a digit level hold K (g + 1) 2^12 < 2^31, i.e. K <= 65536 per CTA; longer
// contractions (the G3 of a wide MPO bond: K = w M) run split-K with at least K / 65536 splits, whose
// partial tiles are summed in FP64 (launch_ozaki_gemm_maps), up to 16 splits.
constexpr double OZ_MAX_K = 16.0 * 65536.0;
static int g_ozaki_slices = 7;
static double g_ozaki_min_work = 4.0e6;

static int gemm_dispatch(cudaStream_t st, int path, int m, int n, int k, const double* A, long lda,
                         const double* B, long ldb, double* C, long ldc) {
  // tensor-core path for contractions large enough to fill 128x128 tiles; boundary sites stay on
  // the exact DMMA kernel
  if (path == 1 && (double)m * n * k >= g_ozaki_min_work && k <= OZ_MAX_K)
    return rn_ozaki_gemm_tn(st, m, n, k, A, lda, B, ldb, C, ldc, g_ozaki_slices);
  return launch_gemm_tn_f64(st, m, n, k, A, lda, B, ldb, C, ldc, 0, 1, 0, 0, 0);
}

// A GEMM operand split into int8 digits, with its TMA descriptor (tcgen05 path).
struct OzOperand {
  signed char* q = nullptr;
  double* scale = nullptr;
  CUtensorMap map;      // 128-row boxes
  CUtensorMap map64;    // 64-row boxes: the B operand of the CTA-pair kernel
  int rows = 0, K = 0, nslices = 0;
};

static int oz_alloc(cudaStream_t st, OzOperand& o, int rows, int K, int nslices) {
  o.rows = rows; o.K = K; o.nslices = nslices;
  RN_CHECK(cudaMallocAsync((void**)&o.q, ozaki_split_bytes(rows, K, nslices) + 16, st));
  RN_CHECK(cudaMallocAsync((void**)&o.scale, sizeof(double) * (size_t)rows, st));
  int err = ozaki_make_map(&o.map, o.q, (long)nslices * rows, (K + 15) & ~15, 128);
  if (err) return err;
  return ozaki_make_map(&o.map64, o.q, (long)nslices * rows, (K + 15) & ~15, 64);
}

static void oz_free(cudaStream_t st, OzOperand& o) {
  if (o.q) cudaFreeAsync(o.q, st);
  if (o.scale) cudaFreeAsync(o.scale, st);
  o.q = nullptr; o.scale = nullptr;
}

static bool use_ozaki(int path, double m, double n, double k) {
  return path == 1 && m * n * k >= g_ozaki_min_work && k <= OZ_MAX_K;
}

// GEMM with pre-split operands (either may be split on the fly from X when `fresh` is given)
static int oz_gemm(cudaStream_t st, OzOperand& a, const double* a_fresh, long lda, OzOperand& b,
                   const double* b_fresh, long ldb, double* C, long ldc, const double* dotv = nullptr,
                   double* dot_partial = nullptr) {
  int err;
  if (a_fresh) { err = launch_ozaki_split(st, a_fresh, lda, a.rows, a.K, a.nslices, a.q, a.scale); if (err) return err; }
  if (b_fresh) { err = launch_ozaki_split(st, b_fresh, ldb, b.rows, b.K, b.nslices, b.q, b.scale); if (err) return err; }
  return launch_ozaki_gemm_maps(st, a.rows, b.rows, a.K, a.nslices, &a.map, a.scale, &b.map, &b.map64, b.scale, C, ldc,
                                dotv, dot_partial);
}

}  // namespace rn

using namespace rn;

struct rn_hop_plan {
  bool oz1 = false, oz3 = false;          // which GEMMs run on the tcgen05 split path
  bool fuse_w = false;                    // last MPO application writes G3's digits directly
  OzOperand ozL, ozC, ozT, ozR;
  int cplx, es, nsite, path;
  int La, Lb, Lc, Rl, Rf, Rk;
  int d1, g1, d2, g2;
  WCsr w1, w2;
  const double* L;    // caller-owned, (La*Lb) x (Lc*es)
  double* Rb;         // packed right operand of G3: (Rl*es) x (Rf*Rk*es); == R for real dtype
  bool own_Rb;
  double *Cb, *T1, *T2, *T3;
  long rest;          // product of the centre indices between c and k (physical and ancilla)
  long launches;
};

static int hop_plan_fill(rn_hop_plan* p, void* stream, int cplx, int nsite,
                                  const void* L, int La, int Lb, int Lc, const void* R, int Rl,
                                  int Rf, int Rk, int d1, int g1, int d2, int g2, int w1_F,
                                  const int* w1_rowptr, const int* w1_pq, const double* w1_val,
                                  int w2_F, const int* w2_rowptr, const int* w2_pq,
                                  const double* w2_val, int path) {
  cudaStream_t st = (cudaStream_t)stream;
  p->cplx = cplx; p->es = cplx ? 2 : 1; p->nsite = nsite; p->path = path;
  p->La = La; p->Lb = Lb; p->Lc = Lc; p->Rl = Rl; p->Rf = Rf; p->Rk = Rk;
  p->d1 = nsite >= 1 ? d1 : 1; p->g1 = nsite >= 1 ? g1 : 1;
  p->d2 = nsite >= 2 ? d2 : 1; p->g2 = nsite >= 2 ? g2 : 1;
  p->L = (const double*)L;
  p->launches = 0;
  const int es = p->es;
  if (nsite >= 1) {
    p->w1.P = Lb; p->w1.Q = p->d1; p->w1.D = p->d1; p->w1.F = w1_F;
    p->w1.rowptr = w1_rowptr; p->w1.pq = w1_pq; p->w1.val = w1_val;
  }
  if (nsite == 2) {
    p->w2.P = w1_F; p->w2.Q = p->d2; p->w2.D = p->d2; p->w2.F = w2_F;
    p->w2.rowptr = w2_rowptr; p->w2.pq = w2_pq; p->w2.val = w2_val;
  }
  const int wlast = nsite == 0 ? Lb : (nsite == 1 ? w1_F : w2_F);
  if (wlast != Rf) return (int)cudaErrorInvalidValue;
  p->rest = (long)p->d1 * p->g1 * p->d2 * p->g2;
  const long n1 = p->rest * Rk;  // columns of T1 (elements)
  p->Cb = p->T1 = p->T2 = p->T3 = nullptr; p->Rb = nullptr; p->own_Rb = false;
  const long rows3 = (long)La * p->d1 * p->g1 * p->d2 * p->g2;
  const long K3 = (long)Rf * Rk * es;
  p->oz1 = use_ozaki(path, (double)La * Lb, (double)n1 * es, (double)Lc * es);
  p->oz3 = use_ozaki(path, (double)rows3, (double)Rl * es, (double)K3);
  RN_CHECK(cudaMallocAsync((void**)&p->T1, sizeof(double) * es * (size_t)La * Lb * n1, st));
  int err;
  if (!p->oz1)   // DMMA path: "math B"[c,(rest,k)] transposed to K-major (and realified) per apply
    RN_CHECK(cudaMallocAsync((void**)&p->Cb, sizeof(double) * es * es * (size_t)n1 * Lc, st));
  if (!p->oz3) {
    if (cplx) {
      RN_CHECK(cudaMallocAsync((void**)&p->Rb, sizeof(double) * 4 * (size_t)Rl * Rf * Rk, st));
      p->own_Rb = true;
      if ((err = launch_pack(st, 1, 1, 0, Rl, Rf * Rk, R, (long)Rf * Rk, 1, p->Rb, (long)Rf * Rk * 2))) return err;
    } else {
      p->Rb = (double*)R;
    }
  }
  // the last MPO application can emit G3's left-operand digits directly when a row fits in smem
  const int lastF = nsite == 2 ? w2_F : w1_F;
  p->fuse_w = p->oz3 && nsite >= 1 && (size_t)lastF * Rk * es * sizeof(double) <= 160 * 1024;
  if (nsite >= 1 && !(nsite == 1 && p->fuse_w))
    RN_CHECK(cudaMallocAsync((void**)&p->T2, sizeof(double) * es * (size_t)La * p->d1 * p->g1 * w1_F * p->d2 * p->g2 * Rk, st));
  if (nsite == 2 && !p->fuse_w)
    RN_CHECK(cudaMallocAsync((void**)&p->T3, sizeof(double) * es * (size_t)La * p->d1 * p->g1 * p->d2 * p->g2 * w2_F * Rk, st));
  if (p->oz1) {
    if ((err = oz_alloc(st, p->ozL, La * Lb, Lc * es, g_ozaki_slices))) return err;
    if ((err = oz_alloc(st, p->ozC, (int)(n1 * es), Lc * es, g_ozaki_slices))) return err;
    if ((err = launch_ozaki_split(st, p->L, (long)Lc * es, La * Lb, Lc * es, g_ozaki_slices, p->ozL.q, p->ozL.scale))) return err;
  }
  if (p->oz3) {
    if ((err = oz_alloc(st, p->ozT, (int)rows3, (int)K3, g_ozaki_slices))) return err;
    if ((err = oz_alloc(st, p->ozR, Rl * es, (int)K3, g_ozaki_slices))) return err;
    if (cplx)
      err = launch_ozaki_split_bform(st, (const double*)R, K3, Rl, Rf * Rk, g_ozaki_slices, 0, p->ozR.q, p->ozR.scale);
    else
      err = launch_ozaki_split(st, (const double*)R, K3, Rl, (int)K3, g_ozaki_slices, p->ozR.q, p->ozR.scale);
    if (err) return err;
  }
  return 0;
}

// A plan that fails half way (allocation, tensor-map encoding, a split launch) is destroyed before the
// error is returned: nothing it allocated from the stream-ordered pool is leaked.
extern "C" int rn_hop_plan_create(rn_hop_plan** out, void* stream, int cplx, int nsite,
                                  const void* L, int La, int Lb, int Lc, const void* R, int Rl,
                                  int Rf, int Rk, int d1, int g1, int d2, int g2, int w1_F,
                                  const int* w1_rowptr, const int* w1_pq, const double* w1_val,
                                  int w2_F, const int* w2_rowptr, const int* w2_pq,
                                  const double* w2_val, int path) {
  if (out == nullptr || nsite < 0 || nsite > 2) return (int)cudaErrorInvalidValue;
  rn_hop_plan* p = new (std::nothrow) rn_hop_plan();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->Cb = p->T1 = p->T2 = p->T3 = nullptr; p->Rb = nullptr; p->own_Rb = false;
  const int err = hop_plan_fill(p, stream, cplx, nsite, L, La, Lb, Lc, R, Rl, Rf, Rk, d1, g1, d2, g2, w1_F,
                                w1_rowptr, w1_pq, w1_val, w2_F, w2_rowptr, w2_pq, w2_val, path);
  if (err) { rn_hop_plan_destroy(p, stream); return err; }
  *out = p;
  return 0;
}

// MPO application number `which` (0: W1 on T1, 1: W2 on T2) as wapply parameters
static void hop_wparams(const rn_hop_plan* p, int which, WApplyParams& w) {
  if (which == 0) {
    const long Yin = (long)p->g1 * p->d2 * p->g2 * p->Rk;   // y = (g1, h, g2, k)
    const long Y2 = (long)p->d2 * p->g2 * p->Rk;
    w.in = p->T1; w.out = p->T2;
    w.X = p->La; w.P = p->w1.P; w.Q = p->w1.Q; w.Y = (int)Yin;
    w.isx = (long)p->w1.P * p->w1.Q * Yin; w.isp = (long)p->w1.Q * Yin; w.isq = Yin; w.isy = 1;
    w.D = p->w1.D; w.F = p->w1.F; w.Y2 = (int)Y2;
    w.osx = (long)p->d1 * p->g1 * p->w1.F * Y2; w.osd = (long)p->g1 * p->w1.F * Y2;
    w.osy1 = (long)p->w1.F * Y2; w.osf = Y2; w.osy2 = 1;
    w.rowptr = p->w1.rowptr; w.ent_pq = p->w1.pq; w.ent_val = p->w1.val;
  } else {
    const long Yin = (long)p->g2 * p->Rk;   // y = (g2, k)
    w.in = p->T2; w.out = p->T3;
    w.X = (int)((long)p->La * p->d1 * p->g1); w.P = p->w2.P; w.Q = p->w2.Q; w.Y = (int)Yin;
    w.isx = (long)p->w2.P * p->w2.Q * Yin; w.isp = (long)p->w2.Q * Yin; w.isq = Yin; w.isy = 1;
    w.D = p->w2.D; w.F = p->w2.F; w.Y2 = p->Rk;
    w.osx = (long)p->d2 * p->g2 * p->w2.F * p->Rk; w.osd = (long)p->g2 * p->w2.F * p->Rk;
    w.osy1 = (long)p->w2.F * p->Rk; w.osf = p->Rk; w.osy2 = 1;
    w.rowptr = p->w2.rowptr; w.ent_pq = p->w2.pq; w.ent_val = p->w2.val;
  }
  w.YT = 0; w.order = 0;
}

// Number of partial sums hop_apply_dot produces (0: the plan's last GEMM is not on the tensor path
// and the caller has to form the inner product itself).
int rn::hop_dot_tiles(const rn_hop_plan* p) {
  if (!p->oz3) return 0;
  const long rows3 = (long)p->La * p->d1 * p->g1 * p->d2 * p->g2;
  return ozaki_gemm_tiles((int)rows3, p->Rl * p->es);
}

extern "C" int rn_hop_apply(rn_hop_plan* p, void* stream, const void* c_in, void* out) {
  return rn::hop_apply_dot(p, (cudaStream_t)stream, c_in, out, nullptr);
}

// H_eff . c_in, optionally with the partial sums of Re <c_in, H_eff c_in> (one pair per output tile
// of the last GEMM) written to dot_partial: the Lanczos alpha without a separate pass.
int rn::hop_apply_dot(rn_hop_plan* p, cudaStream_t st, const void* c_in, void* out, double* dot_partial) {
  const int es = p->es, cplx = p->cplx;
  const long n1 = p->rest * p->Rk;
  int err;
  // G1: T1[(a,b), (rest,k)] = L[(a,b),c] . C[c,(rest,k)]; the right operand is the transposed
  // (and, when complex, realified) view of c_in
  if (p->oz1) {
    err = launch_ozaki_split_t(st, cplx, c_in, n1, (int)n1, p->Lc, p->ozC.nslices, p->ozC.q, p->ozC.scale);
    if (err) return err;
    err = oz_gemm(st, p->ozL, nullptr, 0, p->ozC, nullptr, 0, p->T1, n1 * es);
  } else {
    err = launch_pack(st, cplx, cplx ? 1 : 0, 0, (int)n1, p->Lc, c_in, 1, n1, p->Cb, (long)p->Lc * es);
    if (err) return err;
    err = gemm_dispatch(st, 0, p->La * p->Lb, (int)(n1 * es), p->Lc * es, p->L, (long)p->Lc * es,
                        p->Cb, (long)p->Lc * es, p->T1, n1 * es);
  }
  if (err) return err;
  const double* lastT = p->T1;
  long rows3 = p->La;           // rows of the G3 left operand
  p->launches += 2;
  bool digits_ready = false;    // ozT already holds the digits of G3's left operand
  for (int which = 0; which < p->nsite; ++which) {
    WApplyParams w;
    hop_wparams(p, which, w);
    const bool last = which == p->nsite - 1;
    if (last && p->fuse_w) {
      err = launch_wapply_split(st, cplx, w, p->ozT.nslices, p->ozT.q, p->ozT.scale);
      if (err != 0) return err > 1 ? err : (int)cudaErrorInvalidValue;   // fit was checked at creation
      digits_ready = true;
    } else {
      err = launch_wapply(st, cplx, w);
      if (err) return err;
    }
    lastT = (const double*)w.out;
    rows3 *= (which == 0 ? (long)p->d1 * p->g1 : (long)p->d2 * p->g2);
    p->launches += 1;
  }
  // G3: out[(rows3), l] = T[(rows3), (f,k)] . R[l, (f,k)]
  const long K3 = (long)p->Rf * p->Rk * es;
  if (p->oz3)
    err = oz_gemm(st, p->ozT, digits_ready ? nullptr : lastT, K3, p->ozR, nullptr, 0, (double*)out, (long)p->Rl * es,
                  dot_partial ? (const double*)c_in : nullptr, dot_partial);
  else
    err = gemm_dispatch(st, 0, (int)rows3, p->Rl * es, (int)K3, lastT, K3, p->Rb, K3,
                        (double*)out, (long)p->Rl * es);
  p->launches += 1;
  return err;
}

extern "C" long rn_hop_plan_launches(const rn_hop_plan* p) { return p ? p->launches : 0; }

extern "C" int rn_hop_plan_destroy(rn_hop_plan* p, void* stream) {
  if (!p) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cb) cudaFreeAsync(p->Cb, st);
  if (p->T1) cudaFreeAsync(p->T1, st);
  if (p->T2) cudaFreeAsync(p->T2, st);
  if (p->T3) cudaFreeAsync(p->T3, st);
  if (p->own_Rb && p->Rb) cudaFreeAsync(p->Rb, st);
  oz_free(st, p->ozL); oz_free(st, p->ozC); oz_free(st, p->ozT); oz_free(st, p->ozR);
  delete p;
  return 0;
}

// Environment update.  domain 0 ("L"): env (Ea,Eb,Ec), bra (Ea,d,g,Mf), ket (Ec,d,g,Mh);
// domain 1 ("R"): env (Ea,Eb,Ec), bra (Mf,d,g,Ea), ket (Mh,d,g,Ec).  out (Mf, F, Mh).
// `bra` is the UN-conjugated bra-side site tensor (conjugated inside); W is CSR in the
// orientation documented in rn_b200.h.
extern "C" int rn_env_update(void* stream, int cplx, int domain, const void* env, int Ea, int Eb,
                             int Ec, const void* bra, const void* ket, int d, int g, int Mf,
                             int Mh, int F, const int* rowptr, const int* pq, const double* val,
                             void* out, int path) {
  cudaStream_t st = (cudaStream_t)stream;
  const int es = cplx ? 2 : 1;
  int err;
  double *B1 = nullptr, *X1 = nullptr, *Y = nullptr, *A3 = nullptr, *B3 = nullptr;
  // scratch is released on every exit path (a failed launch must not leak pool memory)
  struct Scratch {
    cudaStream_t st;
    double **p[5];
    ~Scratch() { for (double** q : p) if (*q) cudaFreeAsync(*q, st); }
  } guard{st, {&B1, &X1, &Y, &A3, &B3}};
  const long dg = (long)d * g;
  if (domain == 0) {
    // G1: X1[(a,b),(e,l,h)] = env[(a,b),c] . ket[c,(e,l,h)]
    const long n1 = dg * Mh;
    RN_CHECK(cudaMallocAsync((void**)&B1, sizeof(double) * es * es * (size_t)n1 * Ec, st));
    RN_CHECK(cudaMallocAsync((void**)&X1, sizeof(double) * es * (size_t)Ea * Eb * n1, st));
    err = launch_pack(st, cplx, cplx ? 1 : 0, 0, (int)n1, Ec, ket, 1, n1, B1, (long)Ec * es);
    if (err) return err;
    err = gemm_dispatch(st, path, Ea * Eb, (int)(n1 * es), Ec * es, (const double*)env, (long)Ec * es,
                        B1, (long)Ec * es, X1, n1 * es);
    if (err) return err;
    // W: Y[a, d, l, f, h] = W[b,d,e,f] X1[a,b,e,(l,h)]
    RN_CHECK(cudaMallocAsync((void**)&Y, sizeof(double) * es * (size_t)Ea * dg * F * Mh, st));
    WApplyParams w;
    const long Yin = (long)g * Mh;
    w.in = X1; w.out = Y; w.X = Ea; w.P = Eb; w.Q = d; w.Y = (int)Yin;
    w.isx = (long)Eb * d * Yin; w.isp = (long)d * Yin; w.isq = Yin; w.isy = 1;
    w.D = d; w.F = F; w.Y2 = Mh;
    w.osx = dg * F * Mh; w.osd = (long)g * F * Mh; w.osy1 = (long)F * Mh; w.osf = Mh; w.osy2 = 1;
    w.rowptr = rowptr; w.ent_pq = pq; w.ent_val = val; w.YT = 0; w.order = 0;
    err = launch_wapply(st, cplx, w);
    if (err) return err;
    // G3: out[f,(g,h)] = sum_(a,d,l) conj(bra)[(a,d,l),f] . Y[(a,d,l),(g,h)]
    const long K3 = (long)Ea * dg;
    RN_CHECK(cudaMallocAsync((void**)&A3, sizeof(double) * es * (size_t)Mf * K3, st));
    RN_CHECK(cudaMallocAsync((void**)&B3, sizeof(double) * es * es * (size_t)F * Mh * K3, st));
    err = launch_pack(st, cplx, 0, cplx ? 1 : 0, Mf, (int)K3, bra, 1, Mf, A3, K3 * es);
    if (err) return err;
    err = launch_pack(st, cplx, cplx ? 1 : 0, 0, F * Mh, (int)K3, Y, 1, (long)F * Mh, B3, K3 * es);
    if (err) return err;
    err = gemm_dispatch(st, path, Mf, F * Mh * es, (int)(K3 * es), A3, K3 * es, B3, K3 * es,
                        (double*)out, (long)F * Mh * es);
    if (err) return err;
  } else {
    // G1: X1[(h,e,l),(a,b)] = ket[(h,e,l),c] . env[(a,b),c]
    const long m1 = (long)Mh * dg, n1 = (long)Ea * Eb;
    RN_CHECK(cudaMallocAsync((void**)&X1, sizeof(double) * es * (size_t)m1 * n1, st));
    const double* Bop = (const double*)env;
    if (cplx) {
      RN_CHECK(cudaMallocAsync((void**)&B1, sizeof(double) * 4 * (size_t)n1 * Ec, st));
      err = launch_pack(st, 1, 1, 0, (int)n1, Ec, env, Ec, 1, B1, (long)Ec * 2);
      if (err) return err;
      Bop = B1;
    }
    err = gemm_dispatch(st, path, (int)m1, (int)(n1 * es), Ec * es, (const double*)ket, (long)Ec * es,
                        Bop, (long)Ec * es, X1, n1 * es);
    if (err) return err;
    // W: Yt[f, h, d, (l,a)] = W'[p=b, D=d, q=e, F=f] X1[h, e, (l,a), b]
    const long K3 = dg * Ea;
    RN_CHECK(cudaMallocAsync((void**)&Y, sizeof(double) * es * (size_t)F * Mh * K3, st));
    WApplyParams w;
    const long Yin = (long)g * Ea;
    w.in = X1; w.out = Y; w.X = Mh; w.P = Eb; w.Q = d; w.Y = (int)Yin;
    w.isx = (long)d * Yin * Eb; w.isq = Yin * Eb; w.isy = Eb; w.isp = 1;
    w.D = d; w.F = F; w.Y2 = (int)Yin;
    w.osf = (long)Mh * K3; w.osx = K3; w.osd = Yin; w.osy1 = 0; w.osy2 = 1;
    w.rowptr = rowptr; w.ent_pq = pq; w.ent_val = val; w.YT = 0; w.order = 0;
    err = launch_wapply(st, cplx, w);
    if (err) return err;
    // G3: out[f,(g,h)] = sum_(d,l,a) conj(bra)[f,(d,l,a)] . Yt[(g,h),(d,l,a)]
    const double* Aop = (const double*)bra;
    const double* B3op = Y;
    if (cplx) {
      RN_CHECK(cudaMallocAsync((void**)&A3, sizeof(double) * 2 * (size_t)Mf * K3, st));
      RN_CHECK(cudaMallocAsync((void**)&B3, sizeof(double) * 4 * (size_t)F * Mh * K3, st));
      err = launch_pack(st, 1, 0, 1, Mf, (int)K3, bra, K3, 1, A3, K3 * 2);
      if (err) return err;
      err = launch_pack(st, 1, 1, 0, F * Mh, (int)K3, Y, K3, 1, B3, K3 * 2);
      if (err) return err;
      Aop = A3; B3op = B3;
    }
    err = gemm_dispatch(st, path, Mf, F * Mh * es, (int)(K3 * es), Aop, K3 * es, B3op, K3 * es,
                        (double*)out, (long)F * Mh * es);
    if (err) return err;
  }
  return 0;
}

// ---- plain matrix product ------------------------------------------------------------------------
// out (M x N) = a (M x K) . b (K x N), row-major, real or complex: xp.tensordot / xp.dot of the sweep
// (absorbing a bond matrix into the next site, overlaps, norms).  Large products run on the
// tcgen05 split path: the left operand is split row-wise through its real view, the right one by
// the transposing split (which also realifies it), so no packed FP64 copy is made.
extern "C" int rn_matmul(void* stream, int cplx, int M, int K, int N, const void* a, const void* b,
                         void* out, int path) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 0 || N <= 0) return 0;
  const int es = cplx ? 2 : 1;
  if (K <= 0) { RN_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * es * (size_t)M * N, st)); return 0; }
  int err;
  if (use_ozaki(path, (double)M, (double)N * es, (double)K * es)) {
    OzOperand oa, ob;
    err = oz_alloc(st, oa, M, K * es, g_ozaki_slices);
    if (!err) err = oz_alloc(st, ob, N * es, K * es, g_ozaki_slices);
    if (!err) err = launch_ozaki_split(st, (const double*)a, (long)K * es, M, K * es, oa.nslices, oa.q, oa.scale);
    if (!err) err = launch_ozaki_split_t(st, cplx, b, N, N, K, ob.nslices, ob.q, ob.scale);
    if (!err) err = oz_gemm(st, oa, nullptr, 0, ob, nullptr, 0, (double*)out, (long)N * es);
    oz_free(st, oa); oz_free(st, ob);
    return err;
  }
  double* bp = nullptr;
  RN_CHECK(cudaMallocAsync((void**)&bp, sizeof(double) * es * es * (size_t)N * K, st));
  err = launch_pack(st, cplx, cplx ? 1 : 0, 0, N, K, b, 1, N, bp, (long)K * es);
  if (!err) err = gemm_dispatch(st, 0, M, N * es, K * es, (const double*)a, (long)K * es, bp, (long)K * es,
                                (double*)out, (long)N * es);
  cudaFreeAsync(bp, st);
  return err;
}

extern "C" int rn_set_ozaki(int nslices, double min_work) {
  if (nslices < 1 || nslices > 8) return (int)cudaErrorInvalidValue;
  g_ozaki_slices = nslices;
  if (min_work >= 0) g_ozaki_min_work = min_work;
  return 0;
}

// ---- one fused Lanczos iteration (krylov.py:55-83 of the reference) ----------------------------
//   w = H_eff v_j ; alpha_j = Re<v_j, w> ; w -= alpha_j v_j + beta_{j-1} v_{j-1} ; beta_j = |w| ;
//   v_{j+1} = w / beta_j
// V is the Krylov stack (row j = v_j, `n` elements per row), alpha/beta are device arrays of
// (value, 0) pairs.  One C call per iteration keeps the host off the critical path.
extern "C" int rn_multi_dot(void*, int, long, int, const double*, long, const double*, double*, double*);
extern "C" int rn_lanczos_update(void*, long, double*, const double*, const double*, const double*,
                                 const double*, double*, double*);
extern "C" int rn_scale_inv(void*, long, const double*, const double*, double*);

extern "C" int rn_lanczos_step(rn_hop_plan* plan, void* stream, long n, double* V, int j, double* alpha,
                               double* beta, double* w, double* ws) {
  const int es = plan->es;
  const long ld = n * es;
  double* vj = V + (long)j * ld;
  int err = rn_hop_apply(plan, stream, vj, w);
  if (err) return err;
  err = rn_multi_dot(stream, plan->cplx, n, 1, vj, ld, w, ws, alpha + 2 * j);
  if (err) return err;
  err = rn_lanczos_update(stream, ld, w, vj, j > 0 ? vj - ld : nullptr, alpha + 2 * j,
                          j > 0 ? beta + 2 * (j - 1) : nullptr, ws, beta + 2 * j);
  if (err) return err;
  return rn_scale_inv(stream, ld, w, beta + 2 * j, vj + ld);
}
