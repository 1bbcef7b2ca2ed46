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
//   recip  : MUFU.RCP64H seed + 3 Newton steps; quotients by Markstein correction => the
//            quotient is correctly rounded (IEEE division) for normal operands.
#pragma once
#include <cuda_runtime.h>

namespace ld {

constexpr double kTrigFastMax = 105615.0;

__device__ __forceinline__ double rcp_seed(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}

// 1/d for finite normal d (any sign), error < 1 ulp (correctly rounded in all but rare cases)
__device__ __forceinline__ double recip(double d) {
    double y = rcp_seed(d);
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
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
    const double kMagic = 6755399441055744.0; // 1.5 * 2^52: rint() by addition
    const double q = fma(x, 0.6366197723675814, kMagic);
    const int n = __double2loint(q);
    const double j = q - kMagic;
    double r = fma(j, -1.5707963267948966, x);
    r = fma(j, -6.123233995736766e-17, r);
    r = fma(j, 1.4973849048591698e-33, r);
    const double z = r * r;
    // fdlibm __kernel_sin / __kernel_cos minimax coefficients on [-pi/4, pi/4]
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
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

// exp(x) for any x (NaN -> NaN, large -> +inf, very negative -> 0)
__device__ __forceinline__ double exp_fast(double x) {
    const double kMagic = 6755399441055744.0;
    // clamp so that k stays within the two-step scaling range; exp is saturated out there anyway
    const double xc = fmin(fmax(x, -760.0), 760.0);
    const double q = fma(xc, 1.4426950408889634, kMagic);
    const int k = __double2loint(q);
    const double j = q - kMagic;
    double r = fma(j, -0.6931471805599453, xc);
    r = fma(j, -2.3190468138462996e-17, r);
    double g = fma(r, 2.5100424157005067e-08, 2.7620138719733994e-07);
    g = fma(g, r, 2.7557268378684192e-06);
    g = fma(g, r, 2.480152119021773e-05);
    g = fma(g, r, 0.00019841269863105968);
    g = fma(g, r, 0.0013888888917281794);
    g = fma(g, r, 0.008333333333330051);
    g = fma(g, r, 0.04166666666662399);
    g = fma(g, r, 0.16666666666666669);
    g = fma(g, r, 0.5000000000000001);
    const double e = fma(r * r, g, r) + 1.0;
    const int k1 = k >> 1, k2 = k - k1; // |k| <= 1097 => both halves are normal powers of two
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
    const double y = (e * s1) * s2;
    return x != x ? x : y; // fmin/fmax swallowed a NaN
}

// brain.py:37-39 : 1 / (exp(-beta x) + 1)
__device__ __forceinline__ double sigmoid_fast(double x, double neg_beta) {
    const double d = exp_fast(neg_beta * x) + 1.0;
    // exp overflow: d = +inf -> exactly 0 like the reference (the Newton steps would give NaN)
    const double dd = fmin(d, 1.7976931348623157e308);
    const double y = div_by(1.0, dd, recip(dd));
    return d > 1.7976931348623157e308 ? 0.0 : y;
}

} // namespace ld
