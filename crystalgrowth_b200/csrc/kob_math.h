// kob_math.h — arithmetic building blocks shared by the CUDA kernels (device) and the CPU oracle (host).
//
// Everything here is written so that host (g++ -ffp-contract=off) and device (nvcc, any -fmad setting)
// execute the SAME sequence of correctly-rounded IEEE-754 operations:
//   * rn_add/rn_sub/rn_mul/rn_div map to __fadd_rn/... on the device (never contracted into FMA, never
//     replaced by approximate division) and to the plain operator on the host;
//   * the portable atan/sin/cos below use only those primitives plus integer work.
// The reference calls libm atanf/cosf/sinf (src/Kobayashi.cpp:162-171, :206), which is not portable
// bit-for-bit across libm builds; the portable versions restate the classic Cephes single/double
// precision algorithms (S. Moshier) and are accurate to ~1-2 ulp, i.e. a rounding-level substitute.
#ifndef KOB_MATH_H
#define KOB_MATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define KOB_HD __host__ __device__ __forceinline__
#else
#define KOB_HD inline
#endif

namespace kob {

// ---------------------------------------------------------------------------------------------------
// Correctly rounded, never-contracted primitives.
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
KOB_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
KOB_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
KOB_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
KOB_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
KOB_HD double rn_add(double a, double b) { return __dadd_rn(a, b); }
KOB_HD double rn_sub(double a, double b) { return __dsub_rn(a, b); }
KOB_HD double rn_mul(double a, double b) { return __dmul_rn(a, b); }
KOB_HD double rn_div(double a, double b) { return __ddiv_rn(a, b); }
#else
KOB_HD float rn_add(float a, float b) { return a + b; }
KOB_HD float rn_sub(float a, float b) { return a - b; }
KOB_HD float rn_mul(float a, float b) { return a * b; }
KOB_HD float rn_div(float a, float b) { return a / b; }
KOB_HD double rn_add(double a, double b) { return a + b; }
KOB_HD double rn_sub(double a, double b) { return a - b; }
KOB_HD double rn_mul(double a, double b) { return a * b; }
KOB_HD double rn_div(double a, double b) { return a / b; }
#endif

// The reference's constants.  PI_F is `3.141'5926f` (ext/DXViewer/DXViewer-3.1.0/include/dx12header.h:22),
// one ulp below (float)pi: bit pattern 0x40490FDA.  The dead-band is FLT_EPSILON (src/Kobayashi.cpp:154)
// and stays the *float* epsilon in an FP64-typed build.
constexpr float REF_PI_F = 3.1415926f;
constexpr float REF_DEADBAND = 1.1920928955078125e-7f;  // FLT_EPSILON = 2^-23

// ---------------------------------------------------------------------------------------------------
// Portable single precision atan / sin / cos.
// ---------------------------------------------------------------------------------------------------
KOB_HD float p_atan(float xx) {
    float x = xx;
    bool neg = false;
    if (x < 0.0f) { neg = true; x = -x; }
    float y;
    if (x > 2.414213562373095f) {            // tan(3pi/8)
        y = 1.5707963267948966f;
        x = -rn_div(1.0f, x);
    } else if (x > 0.4142135623730950f) {    // tan(pi/8)
        y = 0.7853981633974483f;
        x = rn_div(rn_sub(x, 1.0f), rn_add(x, 1.0f));
    } else {
        y = 0.0f;
    }
    const float z = rn_mul(x, x);
    float p = rn_sub(rn_mul(8.05374449538e-2f, z), 1.38776856032e-1f);
    p = rn_add(rn_mul(p, z), 1.99777106478e-1f);
    p = rn_sub(rn_mul(p, z), 3.33329491539e-1f);
    p = rn_add(rn_mul(rn_mul(p, z), x), x);
    y = rn_add(y, p);
    return neg ? -y : y;
}

// Cody-Waite reduction by pi/4 octants; valid (and accurate) for |x| < 8192, far beyond j*theta <= 8*2pi.
KOB_HD void p_sincos_core(float xx, float* s_out, float* c_out) {
    float x = xx;
    bool sneg = false;
    if (x < 0.0f) { sneg = true; x = -x; }
    int j = (int)rn_mul(1.27323954473516f, x);   // 4/pi
    float y = (float)j;
    if (j & 1) { j += 1; y = rn_add(y, 1.0f); }
    j &= 7;
    bool cneg = false;
    if (j > 3) { sneg = !sneg; cneg = !cneg; j -= 4; }
    if (j > 1) cneg = !cneg;
    x = rn_sub(rn_sub(rn_sub(x, rn_mul(y, 0.78515625f)), rn_mul(y, 2.4187564849853515625e-4f)),
               rn_mul(y, 3.77489497744594108e-8f));
    const float z = rn_mul(x, x);
    // sin polynomial on [-pi/4, pi/4]
    float ps = rn_add(rn_mul(-1.9515295891e-4f, z), 8.3321608736e-3f);
    ps = rn_sub(rn_mul(ps, z), 1.6666654611e-1f);
    ps = rn_add(rn_mul(rn_mul(ps, z), x), x);
    // cos polynomial on [-pi/4, pi/4]
    float pc = rn_sub(rn_mul(2.443315711809948e-5f, z), 1.388731625493765e-3f);
    pc = rn_add(rn_mul(pc, z), 4.166664568298827e-2f);
    pc = rn_mul(rn_mul(pc, z), z);
    pc = rn_add(rn_sub(pc, rn_mul(0.5f, z)), 1.0f);
    float s, c;
    if (j == 1 || j == 2) { s = pc; c = ps; } else { s = ps; c = pc; }
    *s_out = sneg ? -s : s;
    *c_out = cneg ? -c : c;
}
KOB_HD float p_sin(float x) { float s, c; p_sincos_core(x, &s, &c); return s; }
KOB_HD float p_cos(float x) { float s, c; p_sincos_core(x, &s, &c); return c; }

// ---------------------------------------------------------------------------------------------------
// Portable double precision atan / sin / cos.
// ---------------------------------------------------------------------------------------------------
KOB_HD double p_atan(double xx) {
    double x = xx;
    bool neg = false;
    if (x < 0.0) { neg = true; x = -x; }
    double y;
    int flag = 0;
    if (x > 2.41421356237309504880) {
        y = 1.57079632679489661923;
        flag = 1;
        x = -rn_div(1.0, x);
    } else if (x <= 0.66) {
        y = 0.0;
    } else {
        y = 7.85398163397448309616e-1;
        flag = 2;
        x = rn_div(rn_sub(x, 1.0), rn_add(x, 1.0));
    }
    const double z = rn_mul(x, x);
    double p = -8.750608600031904122785e-1;
    p = rn_add(rn_mul(p, z), -1.615753718733365076637e1);
    p = rn_add(rn_mul(p, z), -7.500855792314704667340e1);
    p = rn_add(rn_mul(p, z), -1.228866684490136173410e2);
    p = rn_add(rn_mul(p, z), -6.485021904942025371773e1);
    double q = rn_add(z, 2.485846490142306297962e1);
    q = rn_add(rn_mul(q, z), 1.650270098316988542046e2);
    q = rn_add(rn_mul(q, z), 4.328810604912902668951e2);
    q = rn_add(rn_mul(q, z), 4.853903996359136964868e2);
    q = rn_add(rn_mul(q, z), 1.945506571482613964425e2);
    double r = rn_div(rn_mul(z, p), q);
    r = rn_add(rn_mul(x, r), x);
    if (flag == 2) r = rn_add(r, 0.5 * 6.123233995736765886130e-17);
    else if (flag == 1) r = rn_add(r, 6.123233995736765886130e-17);
    y = rn_add(y, r);
    return neg ? -y : y;
}

KOB_HD void p_sincos_core(double xx, double* s_out, double* c_out) {
    double x = xx;
    bool sneg = false;
    if (x < 0.0) { sneg = true; x = -x; }
    long long j = (long long)rn_mul(1.27323954473516268615, x);  // 4/pi; |x| < 2^30 assumed
    double y = (double)j;
    if (j & 1) { j += 1; y = rn_add(y, 1.0); }
    j &= 7;
    bool cneg = false;
    if (j > 3) { sneg = !sneg; cneg = !cneg; j -= 4; }
    if (j > 1) cneg = !cneg;
    const double z = rn_sub(rn_sub(rn_sub(x, rn_mul(y, 7.85398125648498535156e-1)),
                                   rn_mul(y, 3.77489470793079817668e-8)),
                            rn_mul(y, 2.69515142907905952645e-15));
    const double zz = rn_mul(z, z);
    double ps = 1.58962301576546568060e-10;
    ps = rn_add(rn_mul(ps, zz), -2.50507477628578072866e-8);
    ps = rn_add(rn_mul(ps, zz), 2.75573136213857245213e-6);
    ps = rn_add(rn_mul(ps, zz), -1.98412698295895385996e-4);
    ps = rn_add(rn_mul(ps, zz), 8.33333333332211858878e-3);
    ps = rn_add(rn_mul(ps, zz), -1.66666666666666307295e-1);
    ps = rn_add(z, rn_mul(rn_mul(z, zz), ps));
    double pc = -1.13585365213876817300e-11;
    pc = rn_add(rn_mul(pc, zz), 2.08757008419747316778e-9);
    pc = rn_add(rn_mul(pc, zz), -2.75573141792967388112e-7);
    pc = rn_add(rn_mul(pc, zz), 2.48015872888517045348e-5);
    pc = rn_add(rn_mul(pc, zz), -1.38888888888730564116e-3);
    pc = rn_add(rn_mul(pc, zz), 4.16666666666665929218e-2);
    pc = rn_add(rn_sub(1.0, rn_mul(0.5, zz)), rn_mul(rn_mul(zz, zz), pc));
    double s, c;
    if (j == 1 || j == 2) { s = pc; c = ps; } else { s = ps; c = pc; }
    *s_out = sneg ? -s : s;
    *c_out = cneg ? -c : c;
}
KOB_HD double p_sin(double x) { double s, c; p_sincos_core(x, &s, &c); return s; }
KOB_HD double p_cos(double x) { double s, c; p_sincos_core(x, &s, &c); return c; }

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3"), counter based.
// ---------------------------------------------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

KOB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// Noise draw r in [0,1) for global cell (i, j) at time step `step`, key = seed.
// One Philox call serves the 4 cells i = 4q..4q+3: counter = (q, j, step_lo, step_hi), word i&3.
// r = (word >> 8) * 2^-24 is exact in float and in double, so both precisions see the same field,
// and it depends only on GLOBAL coordinates, which makes the stream independent of the strip layout.
KOB_HD float noise_from_word(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }
KOB_HD float noise_r(uint64_t seed, uint64_t step, uint32_t i, uint32_t j) {
    const Philox4 p = philox4x32_10(i >> 2, j, (uint32_t)step, (uint32_t)(step >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t k = i & 3u;
    const uint32_t w = k == 0u ? p.w[0] : (k == 1u ? p.w[1] : (k == 2u ? p.w[2] : p.w[3]));
    return noise_from_word(w);
}

}  // namespace kob
#endif  // KOB_MATH_H
