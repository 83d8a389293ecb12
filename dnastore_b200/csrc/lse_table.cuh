// The reference's table-based log_sum_exp (src/logsumexp.h:19-86) on the device, bit for bit.
//
// log_sum_exp_unary divides twice by the table step 1e-4 (`(int)(x / .0001)` and `dx / .0001`,
// logsumexp.h:61-72).  An IEEE fp64 division costs the device a reciprocal seed, ~10 dependent
// fused multiply-adds and a slow-path test; these two divisions were most of the fp64 work of the
// sum-product kernels.  The divisor is a constant, so the correctly rounded quotient is had from
// one multiplication and two fused multiply-adds (Markstein's theorem: if y = RN(1/c) and q is
// within one ulp of x/c, then r = x - c*q is exact in one FMA and RN(q + r*y) = RN(x/c)):
//   c = RN(1e-4), y = RN(1/c) = 10000.0 exactly, y*c - 1 = 0.4316 * 2^-53, so q = RN(x*y) is within
//   0.43 + 0.5 < 1 ulp of x/c.
// tests/test_host.py::test_divide_by_table_step_is_ieee_division checks the same three operations
// against hardware division on the CPU over random and adversarial operands (multiples of the step
// +- a few ulps); the GPU parity tests of the forward and pair-HMM kernels hold the result to the
// oracle's bits.
#pragma once
#include <cuda_runtime.h>

namespace dnab {

__device__ __forceinline__ double divByLseStep(double x) {  // == x / .0001, correctly rounded
  const double c = .0001, y = 10000.0;
  const double q = __dmul_rn(x, y);
  const double r = __fma_rn(-c, q, x);
  return __fma_rn(r, y, q);
}

__device__ __forceinline__ double lseUnary(const double* __restrict__ table, double x) {  // logsumexp.h:52-74
  if (x >= 10 || isnan(x) || isinf(x)) return 0;
  const int n = (int)divByLseStep(x);
  const double dx = x - (n * .0001);
  const double f0 = __ldg(table + n), f1 = __ldg(table + n + 1);
  const double df = f1 - f0;
  return f0 + df * divByLseStep(dx);
}

__device__ __forceinline__ double lse(const double* __restrict__ table, double a, double b) {  // logsumexp.h:34-50
  double mx, diff;
  if (a == b) {
    mx = a;
    diff = 0;
  } else if (a < b) {
    mx = b;
    diff = b - a;
  } else {
    mx = a;
    diff = a - b;
  }
  return mx + lseUnary(table, diff);
}

// The same function without branches (select instead of early return), so that two independent
// evaluations can be interleaved by the instruction scheduler.  Bit-identical to lse(): for a == b finite the
// difference is +0 and the interpolation yields table[0] exactly as lseUnary(0) does; for a == b == -inf the
// difference is NaN, which takes the "return 0" arm just like an infinite difference, and -inf + 0 == -inf +
// log 2; every other case evaluates the same expressions.
__device__ __forceinline__ double lseFlat(const double* __restrict__ table, double a, double b) {
  const bool lt = a < b;
  const double mx = lt ? b : a;
  const double diff = lt ? b - a : a - b;
  const bool far = !(diff < 10);  // >= 10, inf or NaN: logsumexp.h:53-55
  const double x = far ? 0. : diff;
  const int n = (int)divByLseStep(x);
  const double dx = x - (n * .0001);
  const double f0 = __ldg(table + n), f1 = __ldg(table + n + 1);
  const double df = f1 - f0;
  const double r = f0 + df * divByLseStep(dx);
  return mx + (far ? 0. : r);
}

}  // namespace dnab
