// kob_row.cuh — the FAST row update shared by every FP32 roofline kernel (kob_step_fast, kob_far2, kob_step_fast2).
//
// One call of row_full() is what the marching warp does per streamed phi row r: pass 1 of the reference
// (src/Kobayashi.cpp:133-171) for row r-1 and pass 2 (:190-215) for row r-2, for the lane's two adjacent cells as one
// packed float2 (Blackwell FFMA2/FADD2/FMUL2).  Because all kernels call the SAME inlined functions on the same inputs,
// their results are bit-identical by construction (tests/test_fast2.py holds them to that).
//
// The data-dependent part (angle state machine, anisotropy, m(T), noise) is cold_block(): written against a tiny set of
// primitives that also compile for the host (g++ -DKOB_HOST_EMU, tests/cpp/cold_check.cpp), so that the formulas are
// checked on the CPU against the reference's own expressions before any GPU time is spent.
//
// Formulas (all rounding-level substitutes for the reference's, gated by the 1e-6 single-step / 1e-4 window tests):
//   * angle (:154-167): atan(gy/gx) = s pi/4 + atan((gy - s gx)/(gx + s gy)), s = sign(gx gy): the argument is in
//     [-1, 1] for every quadrant, so ONE odd degree-15 minimax polynomial serves without min/max/swap selects; the
//     reference's branch offsets (0, PI_F, 2 PI_F) and s pi/4 are folded into four per-quadrant constants.
//   * anisotropy, integer j = 4, 6 (:170-171): with a = cos 2theta = (gx^2 - gy^2)/|g|^2 and b = sin 2theta / 2 =
//     gx gy/|g|^2 (one MUFU.RCP): cos 4theta = 2a^2 - 1, sin 4theta = 4ab; cos 6theta = a(4a^2 - 3),
//     sin 6theta = b(8a^2 - 2).  Other integer j: (c + i s)^j by repeated squaring; any real j: sincosf of j(theta - theta0).
//   * m(T) (:206): atan(x) = s pi/4 + atan((x - s)/(1 + |x|)), s = sign(x), through the same polynomial.
#ifndef KOB_ROW_CUH
#define KOB_ROW_CUH

#include <stdint.h>

#if defined(KOB_HOST_EMU) && !defined(__CUDACC__)
#include <cmath>
#include <cstring>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#define KOB_RD inline
#else
#include <cuda_runtime.h>
#define KOB_RD __device__ __forceinline__
#endif

namespace kob {

// ---- primitives ------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) || (defined(__CUDACC__) && !defined(KOB_HOST_EMU))
KOB_RD float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
KOB_RD float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
KOB_RD float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
KOB_RD float2 f2sub(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
KOB_RD float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KOB_RD float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KOB_RD uint32_t fbits(float x) { return __float_as_uint(x); }
KOB_RD float bitsf(uint32_t u) { return __uint_as_float(u); }
KOB_RD float fast_cos(float t) { return __cosf(t); }
KOB_RD float fast_sin(float t) { return __sinf(t); }
#define KOB_ANY(p) __any_sync(0xffffffffu, (p))
#else
KOB_RD float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
KOB_RD float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
KOB_RD float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
KOB_RD float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
KOB_RD float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
KOB_RD float rcp_approx(float x) { return 1.0f / x; }
KOB_RD uint32_t fbits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
KOB_RD float bitsf(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
KOB_RD float fast_cos(float t) { return cosf(t); }
KOB_RD float fast_sin(float t) { return sinf(t); }
#define KOB_ANY(p) (p)
#endif
KOB_RD float2 f2(float a) { return make_float2(a, a); }
KOB_RD float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
KOB_RD float2 rcp2(float2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
// |mag| with the sign bit of sgn
KOB_RD float copysign_bits(float mag, float sgn) { return bitsf((fbits(mag) & 0x7fffffffu) | (fbits(sgn) & 0x80000000u)); }

// atan on [-1, 1] for a pair of cells: w + w t Q(t), t = w^2, Q = degree-7 minimax (max abs error 8.5e-8 in FP32,
// 1.4 ulp at pi/4 — the class of CUDA's atanf, at half the issue slots because both cells share every FFMA2).
KOB_RD float2 atan11_2(float2 w) {
    const float2 t = f2mul(w, w);
    float2 p = f2fma(f2(0.002622196450829506f), t, f2(-0.015132336877286434f));
    p = f2fma(p, t, f2(0.04112152010202408f));
    p = f2fma(p, t, f2(-0.0736667588353157f));
    p = f2fma(p, t, f2(0.10573916882276535f));
    p = f2fma(p, t, f2(-0.14185971021652222f));
    p = f2fma(p, t, f2(0.1999039649963379f));
    p = f2fma(p, t, f2(-0.33332985639572144f));
    return f2fma(f2mul(w, t), p, w);
}

// (c + i s)^j, 0 <= j <= 16 (warp-uniform), by repeated squaring
KOB_RD void cpow2_rt(int j, float2 c, float2 s, float2& C, float2& S) {
    float2 rc = f2(1.0f), rs = f2(0.0f);
#pragma unroll
    for (int bit = 4; bit >= 0; --bit) {
        const float2 qc = f2fma(rc, rc, f2neg(f2mul(rs, rs))), qs = f2mul(f2add(rc, rc), rs);
        rc = qc; rs = qs;
        if ((j >> bit) & 1) { const float2 tc = f2fma(rc, c, f2neg(f2mul(rs, s))), ts = f2fma(rc, s, f2mul(rs, c)); rc = tc; rs = ts; }
    }
    C = rc; S = rs;
}

// Loop constants of the data-dependent block (kernel parameters: they sit in the constant bank / uniform registers).
struct ColdK {
    float e;                         // dead-band: FLT_EPSILON (src/Kobayashi.cpp:154)
    float off_c, off_y, off_s;       // re-assigned angle = off_c + off_y sy + off_s sx sy + atan(w): PI_F, -PI_F/2, pi/4 - PI_F/2
                                     //   (Q1 pi/4, Q2 PI_F - pi/4, Q3 PI_F + pi/4, Q4 2 PI_F - pi/4; :160-167)
    float half_pi;                   // 0.5 PI_F (:156-158)
    float cfl_p, sfl_p, cfl_m, sfl_m;// cos / sin (j (+-PI_F/2 - theta0)): anisotropy of a case-A cell (gy > 0 / gy < 0)
    float j_rev, jth0_rev;           // j / 2pi and -j theta0 / 2pi: j (theta - theta0) in revolutions (held angles)
    float ebd, epsbar, neg_ebjd;     // eps = epsbar + (epsbar delta) cos, eps' = ((-epsbar j) delta) sin   (:170-171)
    float eps0, epsd0;               // eps, eps' of a cell holding theta = 0
    float cj0, sj0;                  // cos / sin (j theta0): rotation of the anisotropy axes (extension)
    float neg_gamma, gamma_teq;      // gamma (T_eq - T) = fma(T, -gamma, gamma T_eq)   (:206)
    float aop, m_q;                  // alpha / PI_F and (alpha / PI_F) pi/4
    float noise_a;                   // noise amplitude (extension)
    float aniso, theta0;             // any real j: trig on j (theta - theta0)
    int jmode;                       // run-time integer mode (JM == 0)
};

#if !defined(KOB_HOST_EMU) || defined(__CUDACC__)
__device__ __noinline__ void fast_sincos(float arg, float* s, float* c) { sincosf(arg, s, c); }
#else
static inline void fast_sincos(float arg, float* s, float* c) { *s = sinf(arg); *c = cosf(arg); }
#endif

// cos / sin (j theta) of the direction (ux, uy) (any length), integer j (:170-171), and the rotation by j theta0.
template <int JM, bool ROT>
KOB_RD void aniso_cs(const ColdK& K, float2 ux, float2 uy, float2& Cc, float2& Ss) {
    if (JM == 4 || JM == 6) {
        const float2 gxx = f2mul(ux, ux), gyy = f2mul(uy, uy);
        const float2 inv = rcp2(f2add(gxx, gyy));
        const float2 a = f2mul(f2sub(gxx, gyy), inv), b = f2mul(f2mul(ux, uy), inv);
        const float2 a2 = f2mul(a, a);
        if (JM == 6) {
            Cc = f2mul(a, f2fma(a2, f2(4.0f), f2(-3.0f)));
            Ss = f2mul(b, f2fma(a2, f2(8.0f), f2(-2.0f)));
        } else {
            Cc = f2fma(a2, f2(2.0f), f2(-1.0f));
            Ss = f2mul(f2mul(a, b), f2(4.0f));
        }
    } else {
        const float2 r2 = f2fma(ux, ux, f2mul(uy, uy));
        const float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
        cpow2_rt(K.jmode, f2mul(ux, rinv), f2mul(uy, rinv), Cc, Ss);
        if (ROT) {
            const float2 c2 = f2fma(Cc, f2(K.cj0), f2mul(Ss, f2(K.sj0)));
            const float2 s2 = f2fma(Ss, f2(K.cj0), f2neg(f2mul(Cc, f2(K.sj0))));
            Cc = c2; Ss = s2;
        }
    }
}

// The data-dependent block for the lane's two cells of row r-1 (pass 1) and r-2 (reaction term).
//   in : gx, gy gradient of row r-1; th_old = angle a held cell keeps (0 where re-assigned or not read);
//        asg = cell re-assigns its angle; phi2, tq = phi, T of row r-2; q = phi2 (1 - phi2); rq = noise draw r - 1/2
//   out: An = eps^2, Bn = eps eps' (row r-1); th2 = re-assigned angle (valid where asg); radd = reaction (+ noise) of row r-2
// JM: 4 / 6 = compile-time integer mode, 0 = run-time integer mode (K.jmode in 0..16), -1 = any real j (trig).
template <int JM, bool NOISE, bool ROT, bool GEN>
KOB_RD void cold_block(const ColdK& K, float2 gx, float2 gy, const float (&th_old)[2], const bool (&asg)[2], float2 phi2,
                       float2 tq, float2 q, float2 rq, float2& An, float2& Bn, float2& th2, float2& radd) {
    // ---- re-assigned angle (:154-167) ----
    // With sx, sy = +-1 by the SIGN BITS of gx, gy (gy = -0 counts as negative, consistently everywhere below) and
    // s = sx sy: w = (gy - s gx) / (gx + s gy) in [-1, 1], and the quadrant offsets (Q1 pi/4, Q2 PI_F - pi/4, Q3 PI_F + pi/4,
    // Q4 2 PI_F - pi/4) are the bilinear form PI_F - (PI_F/2) sy + (pi/4 - PI_F/2) s: three packed instructions, no selects.
    {
        const float2 sx = make_float2(copysign_bits(1.0f, gx.x), copysign_bits(1.0f, gx.y));
        const float2 sy = make_float2(copysign_bits(1.0f, gy.x), copysign_bits(1.0f, gy.y));
        const float2 sxy = f2mul(sx, sy);
        const float2 w = f2mul(f2sub(gy, f2mul(gx, sxy)), rcp2(f2fma(gy, sxy, gx)));
        float2 off = f2fma(sy, f2(K.off_y), f2(K.off_c));
        off = f2fma(sxy, f2(K.off_s), off);
        th2 = f2add(off, atan11_2(w));
    }
    // ---- reaction term q ((phi - 1/2) + m(T)) [+ noise] of row r-2 (:206-214) ----
    {
        const float2 xa = f2fma(tq, f2(K.neg_gamma), f2(K.gamma_teq));
        const float2 one = make_float2(copysign_bits(1.0f, xa.x), copysign_bits(1.0f, xa.y));
        const float2 w = f2mul(f2sub(xa, one), rcp2(f2fma(xa, one, f2(1.0f))));
        const float2 m = f2fma(atan11_2(w), f2(K.aop), f2mul(one, f2(K.m_q)));
        float2 t = f2add(f2add(phi2, f2(-0.5f)), m);
        if (NOISE) t = f2fma(rq, f2(K.noise_a), t);
        radd = f2mul(q, t);
    }
    // ---- cos / sin (j (theta - theta0)) (:170-171) from the gradient direction ----
    float2 Cc = f2(1.0f), Ss = f2(0.0f);
    if (JM >= 0) aniso_cs<JM, ROT>(K, gx, gy, Cc, Ss);
    // ---- the cells that do not take their direction from the gradient: dead-band in gx (case A, :154-158: theta =
    // +-PI_F/2, anisotropy by two host constants) and held non-zero angles (cos / sin (j (theta - theta0)) by MUFU on the
    // angle reduced to [-1/2, 1/2] revolutions).  Whole saturated regions consist of such cells, so this is kept cheap. ----
    {
        const bool fl0 = asg[0] && fabsf(gx.x) <= K.e, fl1 = asg[1] && fabsf(gx.y) <= K.e;
        const bool h0 = GEN && th_old[0] != 0.f, h1 = GEN && th_old[1] != 0.f;
        if (KOB_ANY(fl0 || fl1 || h0 || h1)) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool fl = k ? fl1 : fl0, held = k ? h1 : h0;
                const bool neg = (k ? gy.y : gy.x) < 0.f;
                float& thk = k ? th2.y : th2.x;
                float& Ck = k ? Cc.y : Cc.x;
                float& Sk = k ? Ss.y : Ss.x;
                thk = fl ? (neg ? -K.half_pi : K.half_pi) : thk;
                if (JM >= 0) {
                    float u = fmaf(th_old[k], K.j_rev, K.jth0_rev);
                    u = (u - rintf(u)) * 6.28318530717958648f;
                    const float ch = GEN ? fast_cos(u) : 0.f, sh = GEN ? fast_sin(u) : 0.f;
                    Ck = fl ? (neg ? K.cfl_m : K.cfl_p) : (held ? ch : Ck);
                    Sk = fl ? (neg ? K.sfl_m : K.sfl_p) : (held ? sh : Sk);
                }
            }
        }
    }
    if (JM < 0) {                                                                               // any real j: trig on the angle
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float th = asg[k] ? (k ? th2.y : th2.x) : th_old[k];
            if (asg[k] || th != 0.f) {
                float Ck, Sk;
                fast_sincos(K.aniso * (th - K.theta0), &Sk, &Ck);
                if (k) { Cc.y = Ck; Ss.y = Sk; } else { Cc.x = Ck; Ss.x = Sk; }
            }
        }
    }
    float2 ep = f2fma(Cc, f2(K.ebd), f2(K.epsbar));                                             // :170
    float2 ed = f2mul(Ss, f2(K.neg_ebjd));                                                      // :171
    const bool d0 = !asg[0] && !(GEN && th_old[0] != 0.f), d1 = !asg[1] && !(GEN && th_old[1] != 0.f);   // holds theta = 0
    ep = make_float2(d0 ? K.eps0 : ep.x, d1 ? K.eps0 : ep.y);
    ed = make_float2(d0 ? K.epsd0 : ed.x, d1 ? K.epsd0 : ed.y);
    An = f2mul(ep, ep);
    Bn = f2mul(ep, ed);
}

// eps^2 and eps eps' of two cells that HOLD their angle (no re-assignment anywhere in the warp's row): the same operations the
// data-dependent block performs for such cells.
template <int JM, bool ROT>
KOB_RD void held_block(const ColdK& K, const float (&th_old)[2], float2& An, float2& Bn) {
    float2 Cc = f2(1.0f), Ss = f2(0.0f);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float& Ck = k ? Cc.y : Cc.x;
        float& Sk = k ? Ss.y : Ss.x;
        if (JM >= 0) {
            float u = fmaf(th_old[k], K.j_rev, K.jth0_rev);
            u = (u - rintf(u)) * 6.28318530717958648f;
            Ck = fast_cos(u);
            Sk = fast_sin(u);
        } else if (th_old[k] != 0.f) {
            fast_sincos(K.aniso * (th_old[k] - K.theta0), &Sk, &Ck);
        }
    }
    float2 ep = f2fma(Cc, f2(K.ebd), f2(K.epsbar));
    float2 ed = f2mul(Ss, f2(K.neg_ebjd));
    const bool d0 = th_old[0] == 0.f, d1 = th_old[1] == 0.f;                                    // holds theta = 0
    ep = make_float2(d0 ? K.eps0 : ep.x, d1 ? K.eps0 : ep.y);
    ed = make_float2(d0 ? K.epsd0 : ed.x, d1 ? K.epsd0 : ed.y);
    An = f2mul(ep, ep);
    Bn = f2mul(ep, ed);
}

#if defined(__CUDACC__) && !defined(KOB_HOST_EMU)
// ---- the marching warp's register windows and the row update (device only) ------------------------------------------------

// Register windows of one sub-step level; "r" is the phi row consumed in the current iteration.  The depth-2 windows are
// indexed by the PARITY of the iteration instead of being rotated (slot [p] holds the older row, is read, and is then
// overwritten with the new row): a loop over row PAIRS ends with every value in the register it started in, so the row
// loop needs neither full unrolling nor register moves.
struct RowState {
    float2 po[2];               // phi rows r-2 ([p]), r-1 ([p^1])
    float2 gxw[2];              // gx of rows r-2 ([p]), r-1 ([p^1])
    float2 gy2;                 // gy(r-2)
    float2 u1, lp1, lap2;       // u(r-1), c(r-1)+u(r-2), complete 9-point sum of row r-2
    float2 tq1, tu1, tlp1;      // T(r-2), u_T(r-2), c_T(r-2)+u_T(r-3)      [T lags phi by a row]
    float2 A[2], P[2];          // eps^2 and eps*eps'*gx of rows r-3 ([p]), r-2 ([p^1])
    float2 Q2;                  // eps*eps'*gy (r-2)
    uint32_t nxa, nxb;          // Philox words of the odd row, drawn at the even row
    bool have_next;
    __device__ __forceinline__ void clear() {
        po[0] = po[1] = gxw[0] = gxw[1] = gy2 = u1 = lp1 = lap2 = tq1 = tu1 = tlp1 = A[0] = A[1] = P[0] = P[1] = Q2 = make_float2(0.f, 0.f);
        nxa = nxb = 0u; have_next = false;
    }
};

// Loop constants of the stencil part (scalars: packed instructions take them as broadcast operands).
struct RowConst {
    float idx, idy, il, ildt, dtt, K;   // 1/dx, 1/dy, 1/(3 dx dx), dt/(3 dx dx), dt/tau, K
    float A0, B0;                       // eps^2 and eps*eps' of a cell holding theta = 0
};

// T-only row (far field: phi == +0 in the whole footprint): rotates the T windows, returns T+ of row r-2.
__device__ __forceinline__ float2 row_tonly(RowState& S, const RowConst& C, float2 tn, float tw, float te) {
    const float2 thsum = make_float2(tw + tn.y, tn.x + te);
    const float2 tu_new = f2fma(f2(2.0f), tn, thsum);
    const float2 lapt = f2add(S.tlp1, tu_new);
    const float2 nt = f2fma(f2(C.K), f2(0.f), f2fma(lapt, f2(C.ildt), S.tq1));       // :215 with phi+ - phi = +0
    S.tlp1 = f2fma(f2(2.0f), thsum, f2fma(f2(-12.0f), tn, S.tu1));
    S.tu1 = tu_new;
    S.tq1 = tn;
    return nt;
}

// One full row.  Inputs: phi row r (own pair pn, west w, east ee), T row r-1 (tn, tw, te), th_old_in = the angle a cell of
// row r-1 keeps if the state machine holds it (GEN only), draw() = noise draw r - 1/2 for the two cells of row r-2 (called
// only when the data-dependent block runs).  Outputs: phi+/T+ of row r-2; asg / th2 = which cells of row r-1 re-assign their
// angle, and to what (th2 is valid only where asg).  Returns the warp vote "some cell did data-dependent work".
template <int PAR, int JM, bool NOISE, bool ROT, bool GEN, class Draw>
__device__ __forceinline__ bool row_full(RowState& S, const RowConst& C, const ColdK& K, float2 pn, float w, float ee, float2 tn,
                                         float tw, float te, const float (&th_old_in)[2], Draw&& draw, float2& np_, float2& nt_,
                                         float2& th2, bool (&asg)[2]) {
    constexpr int OLD = PAR, NEW = PAR ^ 1;            // window slots: [OLD] = the older row (r-2 / r-3), [NEW] = the newer one
    const float2 po0 = S.po[OLD], gx1 = S.gxw[NEW], gx2 = S.gxw[OLD], A2 = S.A[NEW], A3 = S.A[OLD], P3 = S.P[OLD];
    // horizontal neighbours of the pass-1 products of row r-2, issued early: the shuffle latency hides behind pass 1
    const float A_w = __shfl_up_sync(0xffffffffu, A2.y, 1);
    const float A_e = __shfl_down_sync(0xffffffffu, A2.x, 1);
    const float Q_w = __shfl_up_sync(0xffffffffu, S.Q2.y, 1);
    const float Q_e = __shfl_down_sync(0xffffffffu, S.Q2.x, 1);
    // ---- phi row r: horizontal sums and x-gradient; T row r-1: horizontal sums ----
    const float2 hsum = make_float2(w + pn.y, pn.x + ee);
    const float2 gxn = f2mul(make_float2(pn.y - w, ee - pn.x), f2(C.idx));                    // :139
    const float2 thsum = make_float2(tw + tn.y, tn.x + te);
    // ---- pass 1 for row r-1, far-field values first ----
    const float2 gyn = f2mul(f2sub(pn, po0), f2(C.idy));                                      // :140
    float2 An = f2(C.A0), Pn = f2mul(f2(C.B0), gx1), Qn = f2mul(f2(C.B0), gyn);               // cells holding theta = 0
    const float2 q = f2fma(f2neg(po0), po0, po0);                                             // phi (1 - phi) of row r-2
    float2 radd = f2(0.f);
    asg[0] = (gx1.x < -K.e) || (fabsf(gyn.x) > K.e);                                          // :154-167: theta re-assigned
    asg[1] = (gx1.y < -K.e) || (fabsf(gyn.y) > K.e);
    // Votes: `busy` = some cell re-assigns its angle or has phi (1 - phi) != 0 -> the whole data-dependent block; otherwise,
    // `held` = some cell carries a non-zero held angle (the inside of a saturated region: phi == 1 exactly, gradient in the
    // dead-band) -> only eps(theta) of those cells is evaluated; otherwise the far-field constants above stand.  A cell gets the
    // same bits whichever tier its row takes (the skipped terms are exact zeros), so the tiers — which depend on what the
    // OTHER lanes of the warp hold — never show in the results.
    const bool busy = __any_sync(0xffffffffu, asg[0] || asg[1] || q.x != 0.f || q.y != 0.f);
    bool vote = busy;
    if (busy) {
        float th_old[2];
        th_old[0] = (GEN && !asg[0]) ? th_old_in[0] : 0.f;
        th_old[1] = (GEN && !asg[1]) ? th_old_in[1] : 0.f;
        float2 rq = f2(0.f);
        if (NOISE) rq = draw();
        float2 Bn;
        cold_block<JM, NOISE, ROT, GEN>(K, gx1, gyn, th_old, asg, po0, S.tq1, q, rq, An, Bn, th2, radd);
        Pn = f2mul(Bn, gx1);
        Qn = f2mul(Bn, gyn);
    } else if (GEN) {
        const bool held = __any_sync(0xffffffffu, th_old_in[0] != 0.f || th_old_in[1] != 0.f);
        if (held) {
            float2 Bn;
            held_block<JM, ROT>(K, th_old_in, An, Bn);
            Pn = f2mul(Bn, gx1);
            Qn = f2mul(Bn, gyn);
            vote = true;
        }
    }
    // ---- pass 2 for row r-2 ----
    const float2 dA = make_float2(A2.y - A_w, A_e - A2.x);                                    // :190-192
    const float2 dQ = make_float2(Q_w - S.Q2.y, S.Q2.x - Q_e);                                // term2, :201-203
    const float2 gEx = f2mul(dA, f2(C.idx));
    const float2 gEy = f2mul(f2sub(An, A3), f2(C.idy));                                       // :193-195
    float2 sm = f2fma(f2sub(Pn, P3), f2(C.idy), radd);                                        // term1 (:197-199) + reaction
    sm = f2fma(dQ, f2(C.idx), sm);
    sm = f2fma(A2, f2mul(S.lap2, f2(C.il)), sm);                                              // eps^2 * lap(phi)
    sm = f2fma(gEx, gx2, sm);                                                                 // term3, :204
    sm = f2fma(gEy, S.gy2, sm);
    np_ = f2fma(sm, f2(C.dtt), po0);                                                          // :211
    const float2 tu_new = f2fma(f2(2.0f), tn, thsum);                                         // u_T(r-1)
    const float2 lapt = f2add(S.tlp1, tu_new);                                                // 9-point sum of T at row r-2
    nt_ = f2fma(f2(C.K), f2sub(np_, po0), f2fma(lapt, f2(C.ildt), S.tq1));                    // :215
    // ---- advance the windows: the [OLD] slots take this iteration's rows ----
    S.tlp1 = f2fma(f2(2.0f), thsum, f2fma(f2(-12.0f), tn, S.tu1));
    S.tu1 = tu_new;
    S.tq1 = tn;
    const float2 u_new = f2fma(f2(2.0f), pn, hsum);
    S.lap2 = f2add(S.lp1, u_new);
    S.lp1 = f2fma(f2(2.0f), hsum, f2fma(f2(-12.0f), pn, S.u1));
    S.u1 = u_new;
    S.gy2 = gyn;
    S.gxw[OLD] = gxn;
    S.po[OLD] = pn;
    S.A[OLD] = An;
    S.P[OLD] = Pn;
    S.Q2 = Qn;
    return vote;
}
#endif  // device only

}  // namespace kob
#endif  // KOB_ROW_CUH
