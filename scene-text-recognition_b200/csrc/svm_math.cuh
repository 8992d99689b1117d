// svm_math.cuh -- exp(x) for x <= 0, branch-free, ~1 ulp (the RBF kernel value exp(-gamma d^2), Kernel::k_function, src/svm.cpp:355-356).
// The tensor-core epilogues evaluate 16 of these per thread and iteration; libdevice's exp() carries a range branch per call,
// which keeps the compiler from interleaving the 16 dependency chains (measured: 546 cycles per value, one chain at a time).
//   n = rint(x log2 e) by the 1.5 * 2^52 trick, r = x - n ln2 (hi + lo), e^r by the degree-13 Taylor polynomial
//   (|r| <= 0.3466: truncation 4e-18), scaled by adding n to the exponent field.  x is clamped at -708 (e^x stays normal).
#pragma once

namespace ert {

__device__ __forceinline__ double exp_nonpos(double x)
{
	x = fmax(x, -708.0);
	const double magic = 6755399441055744.0;
	const double t = fma(x, 1.4426950408889634, magic);
	const int n = __double2loint(t);
	const double nf = t - magic;
	double r = fma(nf, -6.93147180369123816490e-01, x);
	r = fma(nf, -1.90821492927058770002e-10, r);
	double p = 1.0 / 6227020800.0;
	p = fma(p, r, 1.0 / 479001600.0);
	p = fma(p, r, 1.0 / 39916800.0);
	p = fma(p, r, 1.0 / 3628800.0);
	p = fma(p, r, 1.0 / 362880.0);
	p = fma(p, r, 1.0 / 40320.0);
	p = fma(p, r, 1.0 / 5040.0);
	p = fma(p, r, 1.0 / 720.0);
	p = fma(p, r, 1.0 / 120.0);
	p = fma(p, r, 1.0 / 24.0);
	p = fma(p, r, 1.0 / 6.0);
	p = fma(p, r, 0.5);
	p = fma(p, r, 1.0);
	p = fma(p, r, 1.0);
	return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

} // namespace ert
