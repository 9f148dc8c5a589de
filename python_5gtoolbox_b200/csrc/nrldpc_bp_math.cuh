// nrldpc_bp_math.cuh -- float64 tanh(q/2) and 2 atanh(x) of the sum-product check-node update
// (py5gphy/ldpc/nr_ldpc_decode.py:158-163: np.tanh(Lq / 2), 2 * np.arctanh(x)), written for the inner loop of bp_qc_kernel.
//
// Why not the device libm: ncu on the kernel with libm's tanh / atanh shows 382 warp instructions per edge and iteration of
// which only 90 are FP64 arithmetic -- 17 % are UMOV pairs that materialise the polynomial coefficients, 15 % IMAD moves,
// 12 % branches / BSSY / BSYNC around the two range branches of each function, taken both ways inside most warps
// (profiles/r2_bp_ncu_summary.md).  The two functions below are branch-free, keep their coefficients in the constant
// bank (LDCU.128: two coefficients per instruction) and run their three divisions (divisors of known range) without the range test and slow-path call of the general one.
// Accuracy: within 3 ulp of the correctly rounded value (tests/test_bp_math.py compiles this header for the host and
// checks 4 x 10^6 points against glibc); the decoder's outputs -- hard decisions, status, iteration counts -- are pinned
// by the reference's goldens either way, intermediate LLRs were never bit-identical to NumPy's (different libm).
// Host-compilable on purpose (no intrinsics outside the #ifdef): the CPU test needs no GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define NRLDPC_BPM_HD __host__ __device__ __forceinline__
#else
#define NRLDPC_BPM_HD static inline
#endif

namespace nrldpc {
namespace bpmath {

// [0..11]  1/13!, 1/12!, ..., 1/2!   : expm1(z) = z + z^2 (1/2 + z/6 + ... + z^11/13!), |z| <= ln2 / 2
// [12..21] 1/21, 1/19, ..., 1/3      : atanh(f) = f + f^3 (1/3 + f^2/5 + ... + f^18/21), |f| <= (sqrt2 - 1) / (sqrt2 + 1)
#define NRLDPC_BPM_COEF                                                                                                  \
    {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, \
     2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664,      \
     0.16666666666666666, 0.5,                                                                                           \
     0.047619047619047616, 0.05263157894736842, 0.058823529411764705, 0.06666666666666667, 0.07692307692307693,          \
     0.09090909090909091, 0.1111111111111111, 0.14285714285714285, 0.2, 0.3333333333333333}
#ifdef __CUDACC__
__constant__ double kCoefDev[22] = NRLDPC_BPM_COEF;
#endif
static const double kCoefHost[22] = NRLDPC_BPM_COEF;
#ifdef __CUDA_ARCH__
#define NRLDPC_BPM_C(i) kCoefDev[i]
#else
#define NRLDPC_BPM_C(i) kCoefHost[i]
#endif

constexpr double kLog2e = 1.4426950408889634, kLn2Hi = 0.6931471805599453, kLn2Lo = 2.3190468138462996e-17;
constexpr double kRound = 6755399441055744.0;  // 1.5 * 2^52: x + kRound has rint(x) in its low mantissa word

NRLDPC_BPM_HD int hi_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
NRLDPC_BPM_HD int lo_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)u;
#endif
}
NRLDPC_BPM_HD double from_words(int hi, int lo)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x;
#endif
}
// single-precision helpers of the exponent estimate in atanh_twice
NRLDPC_BPM_HD float coarse_rcpf(float d)
{
#ifdef __CUDA_ARCH__
    return __frcp_rn(d);
#else
    return 1.0f / d;
#endif
}
NRLDPC_BPM_HD int float_bits(float v)
{
#ifdef __CUDA_ARCH__
    return __float_as_int(v);
#else
    int i; std::memcpy(&i, &v, 4); return i;
#endif
}

// n / d for a normal d whose reciprocal is finite and normal (every divisor below lies in [2^-53, 2^59]): the device
// division's own sequence -- MUFU.RCP64H, two Newton steps, quotient, one residual correction -- without its range
// test and slow-path call, which no operand of this file can take.
NRLDPC_BPM_HD double div_normal(double n, double d)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double q = n * r;
    return fma(fma(-d, q, n), r, q);
#else
    return n / d;
#endif
}

// tanh(q / 2) = expm1(|q|) / (expm1(|q|) + 2), sign of q.  |q| is clamped at 40 (tanh(20) rounds to 1).
NRLDPC_BPM_HD double tanh_half(double q)
{
    const double a = fmin(fabs(q), 40.0);
    const double t = fma(a, kLog2e, kRound);
    const int ni = lo_word(t);          // n = rint(a log2 e), 0 ... 58
    const double n = t - kRound;
    double z = fma(n, -kLn2Hi, a);
    z = fma(n, -kLn2Lo, z);             // a = n ln2 + z
    double p = NRLDPC_BPM_C(0);
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, z, NRLDPC_BPM_C(i));
    const double emz = fma(z * z, p, z);               // expm1(z)
    const double s = from_words((1023 + ni) << 20, 0);  // 2^n
    const double em = fma(s, emz, s - 1.0);            // expm1(a) = 2^n expm1(z) + (2^n - 1)
    return copysign(div_normal(em, em + 2.0), q);
}

// 2 atanh(x) = log w, w = (1 + |x|) / (1 - |x|), sign of x, for |x| < 1.
// With s = 2^e the power of two nearest to w on the log scale, m = w / s lies in [sqrt(1/2), sqrt 2) and
// log w = e ln2 + 2 atanh(f), f = (m - 1) / (m + 1) = ((1 + |x|) - s (1 - |x|)) / ((1 + |x|) + s (1 - |x|)):
// numerator and denominator are one fma each (no rounded intermediate w), one division in all.  Below 1/2, where s is
// 1 or 2, they are taken as (1 - s) + |x| (1 + s) and (1 + s) + |x| (1 - s), exact in 1 +- s: below 0.17 e = 0 and
// f = |x| exactly; from 1/2 up 1 - |x| is exact and the form above is used as it stands.  e only has to be roughly right (it moves f inside the series' range), so it
// comes from a single-precision estimate of w.
NRLDPC_BPM_HD double atanh_twice(double x)
{
    const double a = fabs(x);
    const float wf = (float)(1.0 + a) * coarse_rcpf((float)(1.0 - a)) * 1.41421356f;  // >= 1.41, < 2^56
    const int e = (float_bits(wf) >> 23) - 127;                                       // floor(log2 w + 1/2)
    const double s = from_words((1023 + e) << 20, 0);
    const double c = 1.0 + a, d = 1.0 - a;
    const bool upper = a >= 0.5;
    const double num = upper ? fma(-s, d, c) : fma(a, 1.0 + s, 1.0 - s);
    const double den = upper ? fma(s, d, c) : fma(a, 1.0 - s, 1.0 + s);
    const double f = div_normal(num, den);
    const double g = f * f;
    double p = NRLDPC_BPM_C(12);
#pragma unroll
    for (int i = 13; i < 22; ++i) p = fma(p, g, NRLDPC_BPM_C(i));
    const double ed = (double)e;
    const double f2 = f + f;
    double r = fma(f2 * g, p, ed * kLn2Lo);  // the small terms first
    r += f2;
    r = fma(ed, kLn2Hi, r);
    return copysign(r, x);
}

}  // namespace bpmath
}  // namespace nrldpc
