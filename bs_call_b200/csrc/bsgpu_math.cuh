// bsgpu_math.cuh -- table-driven double-precision log / exp for the arguments the genotype model produces.
//
// The model's transcendental arguments are well behaved: log() sees positive normal numbers in (1e-6, 3), exp() sees
// differences of log-likelihoods in [-45, 0].  The CUDA math library's log/exp spend more than half of their ~90
// instructions on cases that cannot occur here (denormals, infinities, NaN, negative arguments, overflow).  These
// versions do the standard table reduction with one FMA-exact step and a short polynomial:
//
//   log(x):  x = 2^k * z, z in [0.6875, 1.375);  i = top 7 mantissa bits;  r = fma(z, invc[i], -1)  (|r| < 0.0079)
//            log x = k ln2 + logc[i] + log1p(r),  log1p by its degree-8 Taylor polynomial (next term < 1.3e-20)
//   exp(x):  x = (k / 128) ln2 + r, |r| <= ln2 / 256;  exp x = 2^(k >> 7) * 2^((k & 127) / 128) * (1 + r + ... + r^5/120)
//
// Error is below 1 ulp + 1e-17 absolute on those domains (tests/test_cpu.py::test_fast_math_accuracy measures it
// against long double on the host build of the same code: host fma() and device DFMA are both exactly rounded, so the
// host result IS the device result).  The tables are built on the host in long double at bsgpu_init.
#pragma once
#include <cstdint>
#include <cstring>
#if defined(__CUDACC__)
#define BSGPU_HD __host__ __device__ __forceinline__
#else
#define BSGPU_HD inline
#endif
#include <cmath>

namespace bsgpu {

constexpr int kLogN = 128;
constexpr int kExpN = 128;
constexpr uint64_t kLogOff = 0x3fe6000000000000ull;      // 0.6875

struct MathTables {
	double logt[kLogN][2];    // invc, logc = -log(invc)
	double exp2t[kExpN];      // 2^(j/128)
};

BSGPU_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long)u);
#else
	double d; memcpy(&d, &u, 8); return d;
#endif
}
BSGPU_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
	return (uint64_t)__double_as_longlong(d);
#else
	uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

// x positive, normal, finite
BSGPU_HD double fast_log(double x, const MathTables *__restrict__ mt) {
	constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;     // ln2 split: hi has 11 trailing zero bits
	const uint64_t ix = double_to_bits(x);
	const uint64_t tmp = ix - kLogOff;
	const int i = (int)((tmp >> 45) & (kLogN - 1));
	const int64_t k = (int64_t)tmp >> 52;
	const double z = bits_to_double(ix - (tmp & 0xfff0000000000000ull));
	const double invc = mt->logt[i][0], logc = mt->logt[i][1];
	const double r = fma(z, invc, -1.0);
	const double kd = (double)k;
	const double w = fma(kd, Ln2hi, logc);          // exact: |k| < 2^10 and Ln2hi has 11 spare bits, logc small
	// log1p(r) - r = r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8); next term r^9/9 < 1.3e-20
	const double r2 = r * r;
	double p = fma(r, -1.0 / 8.0, 1.0 / 7.0);
	p = fma(r, p, -1.0 / 6.0);
	p = fma(r, p, 1.0 / 5.0);
	p = fma(r, p, -1.0 / 4.0);
	p = fma(r, p, 1.0 / 3.0);
	p = fma(r, p, -0.5);
	const double lo = fma(r2, p, kd * Ln2lo);
	return w + (r + lo);
}

// x in [-700, 0]
BSGPU_HD double fast_exp(double x, const MathTables *__restrict__ mt) {
	constexpr double InvLn2N = 0x1.71547652b82fep0 * kExpN;                              // 128 / ln2
	constexpr double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;   // -(ln2 / 128) split
	constexpr double Shift = 0x1.8p52;
	const double t = fma(x, InvLn2N, Shift);
	const uint64_t ki = double_to_bits(t);
	const double kd = t - Shift;
	double r = fma(kd, NegLn2hiN, x);
	r = fma(kd, NegLn2loN, r);
	const uint64_t idx = ki & (kExpN - 1);
	const uint64_t top = (ki >> 7) << 52;               // (k >> 7) into the exponent field; arithmetic via wraparound
	const double s = bits_to_double(double_to_bits(mt->exp2t[idx]) + top);
	// e^r - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)
	const double r2 = r * r;
	double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
	p = fma(r, p, 1.0 / 6.0);
	p = fma(r, p, 0.5);
	const double q = fma(r2, p, r);
	return fma(s, q, s);
}

inline void build_math_tables(MathTables *mt) {
	for (int i = 0; i < kLogN; i++) {
		uint64_t lo = kLogOff + ((uint64_t)i << 45), hi = kLogOff + ((uint64_t)(i + 1) << 45);
		double zlo, zhi;
		memcpy(&zlo, &lo, 8);
		memcpy(&zhi, &hi, 8);
		// interval i is [zlo, zhi): 80 intervals of width 2^-8 below 1.0, then 48 of width 2^-7 up to 1.375
		const long double c = 0.5L * ((long double)zlo + (long double)zhi);
		const double invc = (double)(1.0L / c);
		mt->logt[i][0] = invc;
		mt->logt[i][1] = (double)(-logl((long double)invc));
	}
	// the interval that starts at 1.0 is centred ON 1: r = z - 1 exactly and logc = 0, so log(1) = 0 and arguments just
	// above 1 (the log-sum-exp of a confident call) keep full relative accuracy.  |r| < 2^-7 there; the polynomial
	// degree above is chosen for that width.
	mt->logt[80][0] = 1.0;
	mt->logt[80][1] = 0.0;
	for (int j = 0; j < kExpN; j++) mt->exp2t[j] = (double)exp2l((long double)j / kExpN);
}

}  // namespace bsgpu
