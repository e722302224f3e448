// Library bring-up and the host-buffer entry points (copy in, run the device path, copy out).
#include "common.cuh"
#include <algorithm>
#include <string.h>
#include "rn_b200.h"
#include "internal.cuh"

#include <vector>

namespace rn { long g_launches = 0; }

extern "C" const char* rn_version(void) { return "rn_b200 1"; }

extern "C" long rn_launch_count(void) { return rn::g_launches; }

extern "C" int rn_init(int device, int* sm_count, int* cc) {
  RN_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  RN_CHECK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc) *cc = prop.major * 10 + prop.minor;
  if (prop.major != 10) {
    fprintf(stderr, "[rn_b200] device %d is sm_%d%d; this library is built for sm_100a only\n",
            device, prop.major, prop.minor);
    return (int)cudaErrorInvalidDevice;
  }
  // keep stream-ordered allocations cached in the pool between sweep sites
  cudaMemPool_t pool;
  RN_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thr = ~0ull;
  RN_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  return 0;
}

namespace {

struct HostCsr {
  std::vector<int> rowptr, pq;
  std::vector<double> val;
};

// orientation 0: W'[p=b, D=up, q=down, F=f] = W[b,up,down,f]; 1: W'[p=f, D=up, q=down, F=b]
HostCsr build_csr(const double* W, int Wb, int d, int Wf, int orientation) {
  HostCsr c;
  const int F = orientation == 0 ? Wf : Wb;
  c.rowptr.assign((size_t)d * F + 1, 0);
  for (int up = 0; up < d; ++up)
    for (int fo = 0; fo < F; ++fo) {
      const int P = orientation == 0 ? Wb : Wf;
      for (int p = 0; p < P; ++p)
        for (int q = 0; q < d; ++q) {
          const int b = orientation == 0 ? p : fo, f = orientation == 0 ? fo : p;
          const double v = W[(((size_t)b * d + up) * d + q) * Wf + f];
          if (v != 0.0) { c.pq.push_back(p * d + q); c.val.push_back(v); }
        }
      c.rowptr[(size_t)up * F + fo + 1] = (int)c.pq.size();
    }
  if (c.pq.empty()) { c.pq.push_back(0); c.val.push_back(0.0); }
  return c;
}

struct DevCsr {
  int *rowptr = nullptr, *pq = nullptr;
  double* val = nullptr;
};

int upload_csr(cudaStream_t st, const HostCsr& h, DevCsr& d) {
  RN_CHECK(cudaMallocAsync((void**)&d.rowptr, sizeof(int) * h.rowptr.size(), st));
  RN_CHECK(cudaMallocAsync((void**)&d.pq, sizeof(int) * h.pq.size(), st));
  RN_CHECK(cudaMallocAsync((void**)&d.val, sizeof(double) * h.val.size(), st));
  RN_CHECK(cudaMemcpyAsync(d.rowptr, h.rowptr.data(), sizeof(int) * h.rowptr.size(), cudaMemcpyHostToDevice, st));
  RN_CHECK(cudaMemcpyAsync(d.pq, h.pq.data(), sizeof(int) * h.pq.size(), cudaMemcpyHostToDevice, st));
  RN_CHECK(cudaMemcpyAsync(d.val, h.val.data(), sizeof(double) * h.val.size(), cudaMemcpyHostToDevice, st));
  return 0;
}

void free_csr(cudaStream_t st, DevCsr& d) {
  if (d.rowptr) cudaFreeAsync(d.rowptr, st);
  if (d.pq) cudaFreeAsync(d.pq, st);
  if (d.val) cudaFreeAsync(d.val, st);
}

int to_device(cudaStream_t st, const void* h, size_t bytes, void** d) {
  RN_CHECK(cudaMallocAsync(d, bytes ? bytes : 16, st));
  if (bytes) RN_CHECK(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st));
  return 0;
}

}  // namespace

extern "C" int rn_hop_apply_host(int cplx, int nsite, const void* L, int La, int Lb, int Lc,
                                 const void* R, int Rl, int Rf, int Rk, int d1, int g1, int d2,
                                 int g2, const double* W1, int w1_F, const double* W2, int w2_F,
                                 const void* c_in, void* out, int path) {
  cudaStream_t st = 0;
  const size_t eb = cplx ? 16 : 8;
  if (nsite < 1) { d1 = g1 = 1; }
  if (nsite < 2) { d2 = g2 = 1; }
  const size_t rest = (size_t)d1 * g1 * d2 * g2;
  void *dL, *dR, *dC, *dO;
  int err;
  if ((err = to_device(st, L, eb * La * Lb * Lc, &dL))) return err;
  if ((err = to_device(st, R, eb * Rl * Rf * Rk, &dR))) return err;
  if ((err = to_device(st, c_in, eb * Lc * rest * Rk, &dC))) return err;
  RN_CHECK(cudaMallocAsync(&dO, eb * La * rest * Rl, st));
  DevCsr c1, c2;
  if (nsite >= 1) { if ((err = upload_csr(st, build_csr(W1, Lb, d1, w1_F, 0), c1))) return err; }
  if (nsite == 2) { if ((err = upload_csr(st, build_csr(W2, w1_F, d2, w2_F, 0), c2))) return err; }
  rn_hop_plan* plan = nullptr;
  err = rn_hop_plan_create(&plan, st, cplx, nsite, dL, La, Lb, Lc, dR, Rl, Rf, Rk, d1, g1, d2, g2,
                           w1_F, c1.rowptr, c1.pq, c1.val, w2_F, c2.rowptr, c2.pq, c2.val, path);
  if (err) return err;
  err = rn_hop_apply(plan, st, dC, dO);
  if (err) return err;
  RN_CHECK(cudaMemcpyAsync(out, dO, eb * La * rest * Rl, cudaMemcpyDeviceToHost, st));
  rn_hop_plan_destroy(plan, st);
  free_csr(st, c1); free_csr(st, c2);
  cudaFreeAsync(dL, st); cudaFreeAsync(dR, st); cudaFreeAsync(dC, st); cudaFreeAsync(dO, st);
  RN_CHECK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int rn_env_update_host(int cplx, int domain, const void* env, int Ea, int Eb, int Ec,
                                  const void* bra, const void* ket, int d, int g, int Mf, int Mh,
                                  const double* W, int Wb, int Wf, void* out, int path) {
  cudaStream_t st = 0;
  const size_t eb = cplx ? 16 : 8;
  void *dE, *dB, *dK, *dO;
  int err;
  const int F = domain == 0 ? Wf : Wb;
  if ((domain == 0 ? Wb : Wf) != Eb) return (int)cudaErrorInvalidValue;
  if ((err = to_device(st, env, eb * Ea * Eb * Ec, &dE))) return err;
  if ((err = to_device(st, bra, eb * (size_t)Ea * d * g * Mf, &dB))) return err;
  if ((err = to_device(st, ket, eb * (size_t)Ec * d * g * Mh, &dK))) return err;
  RN_CHECK(cudaMallocAsync(&dO, eb * (size_t)Mf * F * Mh, st));
  DevCsr c;
  if ((err = upload_csr(st, build_csr(W, Wb, d, Wf, domain), c))) return err;
  err = rn_env_update(st, cplx, domain, dE, Ea, Eb, Ec, dB, dK, d, g, Mf, Mh, F, c.rowptr, c.pq,
                      c.val, dO, path);
  if (err) return err;
  RN_CHECK(cudaMemcpyAsync(out, dO, eb * (size_t)Mf * F * Mh, cudaMemcpyDeviceToHost, st));
  free_csr(st, c);
  cudaFreeAsync(dE, st); cudaFreeAsync(dB, st); cudaFreeAsync(dK, st); cudaFreeAsync(dO, st);
  RN_CHECK(cudaStreamSynchronize(st));
  return 0;
}

// scipy.linalg.svd(a, full_matrices=False) for a HOST matrix (optimized_svd, svd_qn.py:13-49):
// upload, rn_svd, download; the singular values come back sorted (descending) like LAPACK's.
extern "C" int rn_svd_host(int cplx, int m, int n, const void* A, void* U, double* S, void* Vh, int path) {
  if (m <= 0 || n <= 0) return 0;
  cudaStream_t st = 0;
  const size_t eb = cplx ? 16 : 8;
  const int k = m < n ? m : n;
  void *dA, *dU, *dV;
  double* dS;
  int err;
  if ((err = to_device(st, A, eb * (size_t)m * n, &dA))) return err;
  RN_CHECK(cudaMallocAsync(&dU, eb * (size_t)m * k, st));
  RN_CHECK(cudaMallocAsync(&dV, eb * (size_t)k * n, st));
  RN_CHECK(cudaMallocAsync((void**)&dS, sizeof(double) * k, st));
  if (k >= 48) err = rn_svd(st, cplx, m, n, dA, n, dU, k, dS, dV, n, 40, path, nullptr);
  else err = rn_svd_jacobi(st, cplx, m, n, dA, n, dU, k, dS, dV, n, 40, nullptr);
  if (err) return err;
  std::vector<char> hu(eb * (size_t)m * k), hv(eb * (size_t)k * n);
  std::vector<double> hs(k);
  RN_CHECK(cudaMemcpyAsync(hu.data(), dU, hu.size(), cudaMemcpyDeviceToHost, st));
  RN_CHECK(cudaMemcpyAsync(hv.data(), dV, hv.size(), cudaMemcpyDeviceToHost, st));
  RN_CHECK(cudaMemcpyAsync(hs.data(), dS, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
  cudaFreeAsync(dA, st); cudaFreeAsync(dU, st); cudaFreeAsync(dV, st); cudaFreeAsync(dS, st);
  RN_CHECK(cudaStreamSynchronize(st));
  std::vector<int> order(k);
  for (int i = 0; i < k; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hs[a] > hs[b]; });
  char* uo = (char*)U;
  char* vo = (char*)Vh;
  for (int j = 0; j < k; ++j) {
    S[j] = hs[order[j]];
    memcpy(vo + eb * (size_t)j * n, hv.data() + eb * (size_t)order[j] * n, eb * (size_t)n);
    for (int r = 0; r < m; ++r)
      memcpy(uo + eb * ((size_t)r * k + j), hu.data() + eb * ((size_t)r * k + order[j]), eb);
  }
  return 0;
}
