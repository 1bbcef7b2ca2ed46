// ld_math.cuh — branch-free fp64 elementary functions for the episode kernel (sm_100a).
//
// The reference evaluates np.sin/np.cos/np.tan/np.exp and python float division
// (geometry.py:26-40,64-66; brain.py:38).  numpy's own results differ from libm by up to 1 ulp
// depending on the host's SIMD level, so bit-equality is not defined for this part of the path
// (DESIGN.md "parity contract"); what matters is an error of about 1 ulp, like CUDA's libm,
// at the lowest instruction count and with straight-line code so that three angles / four
// sigmoids interleave in one warp's instruction stream.  Measured errors: tests/test_gpu_math.py.
//
//   sincos : 3-term Cody-Waite reduction by pi/2 with FMA (valid for |x| <= kTrigFastMax, the
//            same bound CUDA's libm uses for its fast path) + fdlibm kernel polynomials.
//            Larger arguments must be routed to CUDA's sincos() by the caller.
//   exp    : k = rint(x log2 e), r = x - k ln2 (2-term), 1 + r + r^2 g(r) with a degree-9
//            Chebyshev fit of g (|approximation error| < 0.15 ulp), scaled by 2^k in two halves
//            so that overflow -> +inf and underflow -> 0/subnormal come out of the arithmetic.
//   recip  : MUFU.RCP64H seed + 2 Newton steps; quotients by Markstein correction => the
//            quotient is correctly rounded (IEEE division) for normal operands.
#pragma once
#include <cuda_runtime.h>

namespace ld {

constexpr double kTrigFastMax = 105615.0;
constexpr unsigned kTrigFastMaxHi = 0x40f9c8f0u; // high word of 105615.0 (low word 0)

// Polynomial coefficients live in constant memory: a DFMA can take its addend straight from the
// constant bank, whereas literals are rebuilt with two 32-bit moves per use inside the hot loop.
__constant__ double kSinC[6] = {1.58969099521155010221e-10,  -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                -1.98412698298579493134e-04, 8.33333333332248946124e-03,  -1.66666666666666324348e-01};
__constant__ double kCosC[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09,  -2.75573143513906633035e-07,
                                2.48015872894767294178e-05,  -1.38888888888741095749e-03, 4.16666666666666019037e-02};
__constant__ double kExpC[10] = {2.5100424157005067e-08, 2.7620138719733994e-07, 2.7557268378684192e-06, 2.480152119021773e-05,
                                 0.00019841269863105968, 0.0013888888917281794,  0.008333333333330051,   0.04166666666662399,
                                 0.16666666666666669,    0.5000000000000001};
// {2/pi, -pi/2 head, -pi/2 middle, -pi/2 tail (sign folded), log2(e), -ln2 head, -ln2 tail, rint magic}
__constant__ double kRed[8] = {0.6366197723675814,  -1.5707963267948966,  -6.123233995736766e-17,  1.4973849048591698e-33,
                               1.4426950408889634,  -0.6931471805599453,  -2.3190468138462996e-17, 6755399441055744.0};

__device__ __forceinline__ double rcp_seed(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}

// 1/d for finite normal d (any sign), error < 1 ulp (correctly rounded in all but rare cases)
__device__ __forceinline__ double recip(double d) {
    double y = rcp_seed(d); // ~2^-23 relative
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);       // ~2^-46
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);       // rounding-limited
    return y;
}

// n / d given rc = recip(d): Markstein's final step => RN(n/d) for normal results
__device__ __forceinline__ double div_by(double n, double d, double rc) {
    const double q = n * rc;
    const double rem = fma(-q, d, n);
    return fma(rem, rc, q);
}

// sin and cos of x, |x| <= kTrigFastMax (NaN in -> NaN out)
__device__ __forceinline__ void sincos_fast(double x, double& s, double& c) {
    const double magic = kRed[7]; // 1.5 * 2^52: rint() by addition
    const double q = fma(x, kRed[0], magic);
    const int n = __double2loint(q);
    const double j = q - magic;
    double r = fma(j, kRed[1], x);
    r = fma(j, kRed[2], r);
    r = fma(j, kRed[3], r);
    const double z = r * r;
    // fdlibm __kernel_sin / __kernel_cos minimax coefficients on [-pi/4, pi/4]
    double ps = fma(z, kSinC[0], kSinC[1]);
    ps = fma(ps, z, kSinC[2]);
    ps = fma(ps, z, kSinC[3]);
    ps = fma(ps, z, kSinC[4]);
    ps = fma(ps, z, kSinC[5]);
    double pc = fma(z, kCosC[0], kCosC[1]);
    pc = fma(pc, z, kCosC[2]);
    pc = fma(pc, z, kCosC[3]);
    pc = fma(pc, z, kCosC[4]);
    pc = fma(pc, z, kCosC[5]);
    const double sn = fma(ps, r * z, r);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
    // x = j pi/2 + r : quadrant n mod 4
    double ss = (n & 1) ? cs : sn;
    double cc = (n & 1) ? sn : cs;
    // flip sign bits with integer ops (no fp64 pipe): sin negative for n&2, cos negative for (n+1)&2
    const int shi = __double2hiint(ss) ^ ((n & 2) << 30);
    const int chi = __double2hiint(cc) ^ (((n + 1) & 2) << 30);
    s = __hiloint2double(shi, __double2loint(ss));
    c = __hiloint2double(chi, __double2loint(cc));
}

// exp(x) for any x (large -> +inf, very negative -> 0, NaN -> NaN)
__device__ __forceinline__ double exp_fast(double x) {
    const double magic = kRed[7];
    // clamp so that k stays within the two-step scaling range; exp is saturated out there anyway.
    // Written with compares (not fmin/fmax) so that a NaN stays a NaN, as in the reference where a
    // NaN thrust poisons the state and int(nan) finally raises (drone/drone.py:14).
    // Decided on the high word (integer pipe instead of two fp64 compares): |x| >= 760 (incl. +-inf) -> +-760, NaN kept.
    const int xhi = __double2hiint(x);
    const unsigned xabs = (unsigned)xhi & 0x7fffffffu;
    const bool clamp = xabs >= 0x4087c000u && (xabs < 0x7ff00000u || (xabs == 0x7ff00000u && __double2loint(x) == 0));
    const double xc = clamp ? __hiloint2double((xhi & 0x80000000) | 0x4087c000, 0) : x;
    const double q = fma(xc, kRed[4], magic);
    const int k = __double2loint(q);
    const double j = q - magic;
    double r = fma(j, kRed[5], xc);
    r = fma(j, kRed[6], r);
    double g = fma(r, kExpC[0], kExpC[1]);
#pragma unroll
    for (int i = 2; i < 10; ++i) g = fma(g, r, kExpC[i]);
    const double e = fma(r * r, g, r) + 1.0;
    const int k1 = k >> 1, k2 = k - k1; // |k| <= 1097 => both halves are normal powers of two
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
    return (e * s1) * s2;
}

// brain.py:37-39 : 1 / (exp(-beta x) + 1)
__device__ __forceinline__ double sigmoid_fast(double x, double neg_beta) {
    const double d = exp_fast(neg_beta * x) + 1.0;
    // exp overflow: d = +inf -> exactly 0 like the reference (the Newton steps would give NaN); test on the bits
    const bool overflow = __double2hiint(d) == 0x7ff00000 && __double2loint(d) == 0;
    const double dd = overflow ? 1.7976931348623157e308 : d; // NaN stays NaN
    const double y = div_by(1.0, dd, recip(dd));
    return overflow ? 0.0 : y;
}

} // namespace ld
