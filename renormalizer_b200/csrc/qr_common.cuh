// Scalar helpers shared by the QR kernels: real / complex element algebra.
#pragma once
#include "common.cuh"

namespace rn {

template <bool CPLX>
struct Cx;
template <>
struct Cx<false> {
  using T = double;
  __device__ static T zero() { return 0.0; }
  __device__ static T one() { return 1.0; }
  __device__ static T mul(T a, T b) { return a * b; }
  __device__ static T cmul(T a, T b) { return a * b; }  // conj(a) * b
  __device__ static T sub(T a, T b) { return a - b; }
  __device__ static T conj(T a) { return a; }
  __device__ static double abs2(T a) { return a * a; }
  __device__ static double re(T a) { return a; }
  __device__ static double im(T) { return 0.0; }
  __device__ static T make(double r, double) { return r; }
};
template <>
struct Cx<true> {
  using T = double2;
  __device__ static T zero() { return make_double2(0.0, 0.0); }
  __device__ static T one() { return make_double2(1.0, 0.0); }
  __device__ static T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  __device__ static T cmul(T a, T b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
  __device__ static T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
  __device__ static T conj(T a) { return make_double2(a.x, -a.y); }
  __device__ static double abs2(T a) { return a.x * a.x + a.y * a.y; }
  __device__ static double re(T a) { return a.x; }
  __device__ static double im(T a) { return a.y; }
  __device__ static T make(double r, double i) { return make_double2(r, i); }
};

}  // namespace rn
