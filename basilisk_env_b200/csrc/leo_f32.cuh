// leo_f32.cuh -- mixed-precision variant of the LEO tick (BASELINE config 5: FP32-vs-FP64 accuracy / throughput trade-off).
//
// What changes: the four RK4 stage evaluations of a dynamics tick (gravity, J2, facet drag, wheel back-substitution, MRP
// kinematics), the Sun third-body term, the density, the eclipse cone tests and the panel projection run in FP32.
// What does not: the integrated state is kept and ACCUMULATED in FP64 (x += double(sum of FP32 stage increments)), so the
// representation error of a 7e6 m position in FP32 (0.5 m) never enters the state -- only the relative error of the
// per-tick increment does; clocks and time tags, the battery, the wheel-speed invariants and limit tests, the MRP switch,
// every flight-software pass, the thruster / desat logic and the exact per-stage step around thruster switching stay FP64,
// so every discrete decision is taken by the same code as in the FP64 kernel.
#pragma once
#include "leo_core.cuh"

namespace leo {

struct V3f { float x, y, z; };
LEO_HD V3f mkf(float x, float y, float z) { V3f v; v.x = x; v.y = y; v.z = z; return v; }
LEO_HD V3f tof(V3 p) { return mkf((float)p.x, (float)p.y, (float)p.z); }
LEO_HD V3 tod(V3f p) { return mk((double)p.x, (double)p.y, (double)p.z); }
LEO_HD V3f arrf(const float *p) { return mkf(p[0], p[1], p[2]); }
LEO_HD V3f operator+(V3f p, V3f q) { return mkf(p.x + q.x, p.y + q.y, p.z + q.z); }
LEO_HD V3f operator-(V3f p, V3f q) { return mkf(p.x - q.x, p.y - q.y, p.z - q.z); }
LEO_HD V3f operator*(V3f p, float s) { return mkf(p.x * s, p.y * s, p.z * s); }
LEO_HD float dot(V3f p, V3f q) { return p.x * q.x + p.y * q.y + p.z * q.z; }
LEO_HD V3f cross(V3f p, V3f q) { return mkf(p.y * q.z - p.z * q.y, p.z * q.x - p.x * q.z, p.x * q.y - p.y * q.x); }
LEO_HD V3f mv9f(const float *m, V3f v)
{
    return mkf(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
LEO_HD float rsqf(float x)
{
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
struct MrpRotF { float a8, b4, oms2; };
LEO_HD MrpRotF mrp_rotf(V3f s)
{
    float s2 = dot(s, s), den = 1.f + s2, inv = 1.f / (den * den);
    MrpRotF m; m.oms2 = 1.f - s2; m.a8 = 8.f * inv; m.b4 = 4.f * m.oms2 * inv;
    return m;
}
LEO_HD V3f rot_BNf(const MrpRotF &m, V3f s, V3f x) { V3f t = cross(s, x), uu = cross(s, t); return x + uu * m.a8 - t * m.b4; }
LEO_HD V3f rot_NBf(const MrpRotF &m, V3f s, V3f x) { V3f t = cross(s, x), uu = cross(s, t); return x + uu * m.a8 + t * m.b4; }

struct DynF { V3f r, v, s, w; };
// FP32 shadows of what a stage evaluation reads besides the state (see StageIn)
struct StageInF { V3f Lc, gsun, HB, tau_u; float rho, h; };

template <bool J2, bool DIAG>
LEO_HD void eomf(const LeoParamsF &P, const DynF &x, DynF &k, const StageInF &a, float ct, bool thr_on, V3f Fm)
{
    V3f g;
    {
        float ir = rsqf(dot(x.r, x.r));
        float ir3 = ir * ir * ir;
        g = a.gsun + x.r * (-P.mu_c * ir3);
        if (J2) {
            float ir2 = ir * ir, z2 = 5.f * x.r.z * x.r.z * ir2, kk = -P.j2k * ir3 * ir2;
            g = g + mkf(kk * x.r.x * (1.f - z2), kk * x.r.y * (1.f - z2), kk * x.r.z * (3.f - z2));
        }
    }
    MrpRotF R = mrp_rotf(x.s);
    V3f vB = rot_BNf(R, x.s, x.v);
    V3f Mp;
    float Sp;
    {
        float ax = fabsf(vB.x), ay = fabsf(vB.y), az = fabsf(vB.z);
        Sp = P.dragKa[0] * ax + P.dragKa[1] * ay + P.dragKa[2] * az + P.dragKd[0] * vB.x + P.dragKd[1] * vB.y + P.dragKd[2] * vB.z;
        if (DIAG) {
            Mp = mkf(P.dragMa[0][0] * ax + P.dragMd[0][0] * vB.x, P.dragMa[1][1] * ay + P.dragMd[1][1] * vB.y,
                     P.dragMa[2][2] * az + P.dragMd[2][2] * vB.z);
        } else {
            Mp = arrf(P.dragMa[0]) * ax + arrf(P.dragMa[1]) * ay + arrf(P.dragMa[2]) * az
               + arrf(P.dragMd[0]) * vB.x + arrf(P.dragMd[1]) * vB.y + arrf(P.dragMd[2]) * vB.z;
        }
    }
    const float mrho = -a.rho;
    k.v = g + x.v * (mrho * Sp);
    if (thr_on) k.v = k.v + rot_NBf(R, x.s, Fm);
    k.r = x.v;
    V3f rot = a.Lc + cross(Mp * mrho, vB);
    const V3f hw = a.HB + a.tau_u * ct;
    if (DIAG) {
        V3f h = mkf(P.D[0] * x.w.x + hw.x, P.D[4] * x.w.y + hw.y, P.D[8] * x.w.z + hw.z);
        rot = rot - cross(x.w, h);
        k.w = mkf(rot.x * P.Dinv[0], rot.y * P.Dinv[4], rot.z * P.Dinv[8]);
    } else {
        V3f h = mv9f(P.D, x.w) + hw;
        rot = rot - cross(x.w, h);
        k.w = mv9f(P.Dinv, rot);
    }
    {
        float sw = dot(x.s, x.w);
        k.s = x.w * (0.25f * R.oms2) + cross(x.s, x.w) * 0.5f + x.s * (0.5f * sw);
    }
}

// One RK4 step: FP32 stages from the FP32 image of the FP64 state, FP64 accumulation of the weighted slope sum.
template <bool J2, bool DIAG>
LEO_HD Dyn rk4_step_mixed(const LeoParamsF &P, const Dyn &x, const StageInF &a, bool thr_on, V3f Fm)
{
    const float h = a.h, hh = 0.5f * h, h6 = h * (1.0f / 6.0f), h3 = h * (1.0f / 3.0f);
    DynF xf, k, acc;
    xf.r = tof(x.r); xf.v = tof(x.v); xf.s = tof(x.s); xf.w = tof(x.w);
    k.r = k.v = k.s = k.w = acc.r = acc.v = acc.s = acc.w = mkf(0.f, 0.f, 0.f);
    float c = 0.f;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int st = 0; st < 4; st++) {
        DynF xs;
        xs.r = xf.r + k.r * c; xs.v = xf.v + k.v * c; xs.s = xf.s + k.s * c; xs.w = xf.w + k.w * c;
        eomf<J2, DIAG>(P, xs, k, a, c, thr_on, Fm);
        const float wo = (st == 0 || st == 3) ? h6 : h3;
        acc.r = acc.r + k.r * wo; acc.v = acc.v + k.v * wo; acc.s = acc.s + k.s * wo; acc.w = acc.w + k.w * wo;
        c = (st == 2) ? h : hh;
    }
    Dyn xo;
    xo.r = x.r + tod(acc.r); xo.v = x.v + tod(acc.v); xo.s = x.s + tod(acc.s); xo.w = x.w + tod(acc.w);
    return xo;
}

// Sun third-body acceleration in FP32
LEO_HD V3f sun_accelf(const LeoParamsF &P, V3f rs, V3f r)
{
    V3f d = r - rs;
    float is = rsqf(dot(rs, rs)), id = rsqf(dot(d, d));
    return rs * (-P.mu_sun * (is * is * is)) + d * (-P.mu_sun * (id * id * id));
}

// eclipse cone tests in FP32 with a relative guard band of 1e-5 (FP32 resolves the cone radius to ~0.3 m in 6.4e6 m).
// Returns 1 (full sun), 0 (umbra) or -1: inside the band around the cone surfaces or in the penumbra, where the caller runs
// the FP64 evaluation on the FP64 state -- umbra / penumbra / sun stay the same function of position.
#define ECL_BAND_F 1e-5f
LEO_HD float eclipse_coref(const LeoParamsF &PF, const float (&ec)[6], V3f sun_r, V3f r, float s2, float hb2)
{
    const float hp2 = ec[0], inv_hp = ec[1], c1off = ec[2], c2off = ec[3], tan1 = ec[4], tan2 = ec[5];
    const float s0 = -dot(r, sun_r) * inv_hp;
    const float c1 = s0 + c1off, c2 = s0 - c2off;
    const float l2sq = s2 - s0 * s0;
    const float l1 = c1 * tan1, l2 = c2 * tan2;
    const float p2 = l1 * l1, u2 = l2 * l2;
    const bool lit = (hb2 < hp2 * (1.f - ECL_BAND_F)) || (l2sq > p2 * (1.f + ECL_BAND_F) && l2sq > u2 * (1.f + ECL_BAND_F));
    const bool dark = l2sq < u2 * (1.f - ECL_BAND_F) && c2 < 0.f && PF.R_sun > PF.R_planet;
    return lit ? 1.f : (dark ? 0.f : -1.f);
}

}  // namespace leo
