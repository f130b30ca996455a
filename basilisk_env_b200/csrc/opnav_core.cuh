// opnav_core.cuh -- per-environment body of the fused opNav decision step.
//
// One call of opnav_step_env() replaces one `scenario_OpNav.run_sim(action)`
// (/root/reference/basilisk_env/simulators/opNavSimulator.py:225-299, cited ONS:line) plus the bookkeeping of
// `opNavEnv.step` (/root/reference/basilisk_env/envs/opNavEnvironment.py:55-125, ONE:line) for ONE spacecraft:
// 3000 ticks of  [wheel latch -> CSS -> eclipse -> RK4 hub + 4 wheels -> Sun -> simple_nav -> camera]  (dynamics
// process, opNav_models/BSK_OpNavDynamics.py:100-110, OND:line) followed by  [guidance -> MRP feedback ->
// wheel torque map -> circle -> pixelLine -> relative-OD filter]  (FSW process, opNav_models/BSK_OpNavFsw.py,
// ONF:line).  The Vizard / OpenCV image path is replaced by a pinhole projection of the Mars disc.
//
// The filter is the square-root unscented Kalman filter of relativeODuKF restated for a streaming register
// implementation.  With d_i = Y_i - Y_0 the propagated deviations from the central sigma point and
// m = sum_i w d_i, the SR-UKF covariance  sum_i wC_i (Y_i - xBar)(Y_i - xBar)^T + Q  equals
//     Q + sum_{i>=1} w d_i d_i^T + (2 - alpha^2) m m^T
// (all terms positive: no down-date, no stored sigma points).  The 21 entries of that sum are independent FMA
// chains over the sigma points; one 6 x 6 Cholesky per tick gives the factor the next sigma points need.  The
// measurement is linear (y = r), so Pxy and Pyy are blocks of the a-priori covariance and the update is three
// hyperbolic rank-one sweeps.  The oracle keeps Basilisk's formulation (Householder QR + Gill-Murray
// down-dates); the two agree to rounding (tests/test_opnav_*).
//
// Everything is __host__ __device__ (tests/hostcore compiles it with g++); the product only runs it on the GPU.
#pragma once
#include <type_traits>
#include "leo_core.cuh"
#include "opnav_params.h"

namespace opnav {
using namespace leo;

#define ON_HD LEO_HD
#define ON_HD_NOINLINE LEO_HD_NOINLINE

struct Truth { V3 r, v, s, w; double Om[ON_NRW]; };

ON_HD void sincos_hd(double x, double &s, double &c)
{
#ifdef __CUDA_ARCH__
    sincos(x, &s, &c);
#else
    s = sin(x); c = cos(x);
#endif
}

// ------------------------------------------------------------------------------------------------
// noise streams: Philox4x32-10, counter (env_lo, env_hi, tick, stream<<16 | block), key (seed_lo, seed_hi ^ episode)
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// Polynomial coefficients of the device-side noise transforms live in constant memory: an FP64 instruction takes a
// constant-bank operand for free, whereas a 64-bit literal costs two uniform-register moves (two issue slots) per use.
__constant__ double ON_K[40] = {
    // [0..9] atanh series 1/3 .. 1/21
    1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0,
    // [10..16] sin Taylor -1/3! .. -1/15!
    -1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0, -1.0 / 1307674368000.0,
    // [17..24] cos Taylor -1/2! .. 1/16!
    -0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, -1.0 / 87178291200.0, 1.0 / 20922789888000.0,
    // [25..27] sqrt2, -2 ln2, (pi/4) / 2^29
    1.4142135623730951, -2.0 * 0.6931471805599453, 0.78539816339744831 / 536870912.0,
    // [28..39] exp Taylor 1/2! .. 1/13!
    0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0,
    1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};
#endif
#ifdef __CUDA_ARCH__
// -2 ln((x + 1) / 2^32) for a 32-bit draw x: the argument is an integer times a power of two, so the exponent and a
// mantissa in [1/sqrt2, sqrt2) come from integer operations and ln m = 2 atanh((m - 1)/(m + 1)) needs eleven odd terms
// (|s| <= 0.1716: truncation 1e-18).  Agrees with libm's log to ~1e-16 relative.
__device__ __forceinline__ double neg2_log_u32(uint32_t x)
{
    const uint32_t n = x + 1u;                               // 0 stands for 2^32, i.e. u = 1
    if (n == 0u) return 0.0;
    const int lz = __clz((int)n);
    double m = (double)(n << lz) * (1.0 / 2147483648.0);     // [1, 2)
    int e = -1 - lz;
    if (m > ON_K[25]) { m *= 0.5; e += 1; }
    const double sq = (m - 1.0) * frcp(m + 1.0), s2 = sq * sq;
    // degree-9 polynomial in s^2, Estrin form: dependency depth 4 instead of 10 (the walk is a chain of short serial
    // sections; with two warps per SM sub-partition the depth, not the operation count, is what costs)
    const double z2 = s2 * s2, z4 = z2 * z2, z8 = z4 * z4;
    const double a0 = fma(s2, ON_K[1], ON_K[0]), a1 = fma(s2, ON_K[3], ON_K[2]), a2 = fma(s2, ON_K[5], ON_K[4]);
    const double a3 = fma(s2, ON_K[7], ON_K[6]), a4 = fma(s2, ON_K[9], ON_K[8]);
    const double b0 = fma(a1, z2, a0), b1 = fma(a3, z2, a2);
    const double p = fma(a4, z8, fma(b1, z4, b0));
    const double lnm = fma(sq * s2, p, sq);                  // atanh(s)
    return fma((double)e, ON_K[26], -4.0 * lnm);
}
// sin and cos of 2 pi j / 2^32: the octant comes from the top three bits, the remainder (reflected in odd octants, exactly,
// on the integer) is an angle in [0, pi/4] for two short Taylor polynomials (truncation < 1e-17).
__device__ __forceinline__ void sincos_2pi_u32(uint32_t j, double &sn, double &cs)
{
    const uint32_t oct = j >> 29;
    uint32_t fr = j & 0x1FFFFFFFu;
    if (oct & 1u) fr = 0x20000000u - fr;
    const double a = (double)fr * ON_K[27], a2 = a * a;
    // Estrin form of both polynomials (degree 6 and 7 in a^2): depth 4 instead of 7 / 8
    const double a4 = a2 * a2, a8 = a4 * a4;
    const double s0 = fma(a2, ON_K[11], ON_K[10]), s1 = fma(a2, ON_K[13], ON_K[12]), s2_ = fma(a2, ON_K[15], ON_K[14]);
    const double ps = fma(fma(ON_K[16], a4, s2_), a8, fma(s1, a4, s0));
    const double si = fma(a * a2, ps, a);
    const double c0 = fma(a2, ON_K[18], ON_K[17]), c1 = fma(a2, ON_K[20], ON_K[19]), c2 = fma(a2, ON_K[22], ON_K[21]);
    const double c3 = fma(a2, ON_K[24], ON_K[23]);
    const double pc = fma(fma(c3, a4, c2), a8, fma(c1, a4, c0));
    const double co = fma(a2, pc, 1.0);
    const bool swap = ((oct + 1u) & 2u) != 0u;               // octants 1, 2, 5, 6
    const double sv = swap ? co : si, cv = swap ? si : co;
    sn = (oct & 4u) ? -sv : sv;                              // octants 4..7
    cs = ((oct + 2u) & 4u) ? -cv : cv;                       // octants 2..5
}
#endif
// Inline form (the caller's `out` stays in registers: used in the rolled loop of the noise walk, where the code exists once
// anyway); normals4() below is the out-of-line form for the rare pixel-noise draw.
ON_HD void normals4_inl(const OpNavParams &P, int64_t env, int64_t episode, uint32_t tick, uint32_t stream, uint32_t block,
                        double (&out)[4])
{
    uint32_t x[4];
    philox4x32((uint32_t)(uint64_t)env, (uint32_t)((uint64_t)env >> 32), tick, (stream << 16) | block, (uint32_t)P.seed,
               (uint32_t)(P.seed >> 32) ^ (uint32_t)episode, x);
#pragma unroll
    for (int p = 0; p < 2; p++) {
        double s, c;
#ifdef __CUDA_ARCH__
        const double t = neg2_log_u32(x[2 * p]);
        const double rr = t > 0.0 ? t * rsq(t) : 0.0;      // sqrt without the special-operand path
        sincos_2pi_u32(x[2 * p + 1], s, c);
#else
        double u1 = ((double)x[2 * p] + 1.0) * (1.0 / 4294967296.0);
        double u2 = (double)x[2 * p + 1] * (1.0 / 4294967296.0);
        const double rr = sqrt(-2.0 * log(u1));
        s = sin(2.0 * 3.14159265358979323846 * u2); c = cos(2.0 * 3.14159265358979323846 * u2);
#endif
        out[2 * p] = rr * c;
        out[2 * p + 1] = rr * s;
    }
}
ON_HD_NOINLINE void normals4(const OpNavParams &P, int64_t env, int64_t episode, uint32_t tick, uint32_t stream, uint32_t block,
                             double (&out)[4])
{
    normals4_inl(P, env, episode, tick, stream, block, out);
}

// ------------------------------------------------------------------------------------------------
// Sun relative to the Mars barycentre (analytic stand-in for SPICE, zeroBase "mars barycenter", OND:396-401):
// evaluated exactly at four nodes of a decision interval, cubic Lagrange in between.
// ------------------------------------------------------------------------------------------------
struct SunState { V3 r, v; };
ON_HD_NOINLINE SunState sun_from_mars(const OpNavParams &P, double t)
{
    if (P.eph_sun.nseg > 0) {          // ephemeris table instead of the analytic series
        SunState o;
        leo::cheb_eval(P.eph_sun, t, o.r, o.v);
        return o;
    }
    const double PI = 3.14159265358979323846, D2R = PI / 180.0, AUm = 149597870700.0;
    double days = P.epoch_days + t / 86400.0;
    double T = days / 36525.0;
    double a = (1.52371034 + 0.00001847 * T) * AUm, e = 0.09339410 + 0.00007882 * T;
    double I = (1.84969142 - 0.00813131 * T) * D2R, L = (-4.55343205 + 19140.30268499 * T) * D2R;
    double wbar = (-23.94362959 + 0.44441088 * T) * D2R, Om = (49.55953891 - 0.29257343 * T) * D2R;
    double w = wbar - Om, M = L - wbar;
    double n = (19140.30268499 - 0.44441088) * D2R / (36525.0 * 86400.0);
    M = fmod(M, 2.0 * PI);
    double E = M + e * sin(M);
    for (int it = 0; it < 8; it++) E = E - (E - e * sin(E) - M) / (1.0 - e * cos(E));
    double b = a * sqrt(1.0 - e * e);
    double xp = a * (cos(E) - e), yp = b * sin(E);
    double Edot = n / (1.0 - e * cos(E));
    double xd = -a * sin(E) * Edot, yd = b * cos(E) * Edot;
    double cw = cos(w), sw = sin(w), cO = cos(Om), sO = sin(Om), cI = cos(I), sI = sin(I);
    V3 P1 = mk(cw * cO - sw * sO * cI, cw * sO + sw * cO * cI, sw * sI);
    V3 P2 = mk(-sw * cO - cw * sO * cI, -sw * sO + cw * cO * cI, cw * sI);
    double eps = 23.43928 * D2R, ce = cos(eps), se = sin(eps);
    V3 re = P1 * xp + P2 * yp, ve = P1 * xd + P2 * yd;
    SunState o;
    o.r = mk(-re.x, -(ce * re.y - se * re.z), -(se * re.y + ce * re.z));
    o.v = mk(-ve.x, -(ce * ve.y - se * ve.z), -(se * ve.y + ce * ve.z));
    return o;
}
// Four exact evaluations per decision interval (nodes at 0, 1/3, 2/3, 1 of the span), cubic Lagrange in between:
// interpolation error < 1e-5 m on 2.3e11 m; only the Sun POSITION is used on this path.
ON_HD V3 sun_at(const volatile double *n, double t0, double T, double t)
{
    double q = 3.0 * ((t - t0) / T);
    double a = q - 1.0, b = q - 2.0, c = q - 3.0;
    double l0 = -(a * b * c) * (1.0 / 6.0), l1 = (q * b * c) * 0.5, l2 = -(q * a * c) * 0.5, l3 = (q * a * b) * (1.0 / 6.0);
    return mk(n[0], n[1], n[2]) * l0 + mk(n[3], n[4], n[5]) * l1 + mk(n[6], n[7], n[8]) * l2 + mk(n[9], n[10], n[11]) * l3;
}

// ------------------------------------------------------------------------------------------------
// truth dynamics: hub + four balanced wheels, Mars point mass (OND:382-391); thrusters never commanded,
// extForceTorque zero
// ------------------------------------------------------------------------------------------------
// In this scenario the translational and the rotational equations do not couple (point-mass gravity, no drag, no gravity
// gradient torque), so the 16-state RK4 step of SpacecraftPlus is evaluated as two independent RK4 steps -- the same
// arithmetic on every state, with 40 instead of 64 doubles live at the peak.
ON_HD void rk4_translation(const OpNavParams &P, V3 &r, V3 &v, double h)
{
    const V3 r0 = r, v0 = v;
    V3 ro = r0, vo = v0, rs = r0, vs = v0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int st = 0; st < 4; st++) {
        const double ir = rsq(dot(rs, rs));
        const V3 kr = vs, kv = rs * (-P.mu_dyn * (ir * ir * ir));
        const double cw = (st == 0 || st == 3) ? h / 6.0 : h / 3.0;
        const double cn = st == 2 ? h : 0.5 * h;
        ro = ro + kr * cw; vo = vo + kv * cw;
        rs = r0 + kr * cn; vs = v0 + kv * cn;
    }
    r = ro; v = vo;
}
struct Rot { V3 s, w; double Om[ON_NRW]; };
ON_HD void eom_rotation(const OpNavParams &P, const Rot &x, const double (&u)[ON_NRW], Rot &k)
{ // hub + four balanced wheels (back-substitution with the constant matrix D = I - sum Js g g^T), MRP kinematics
    V3 hw = mk(0., 0., 0.), gu = mk(0., 0., 0.);
#pragma unroll
    for (int i = 0; i < ON_NRW; i++) {
        V3 g = arr(P.gs[i]);
        hw = hw + g * (P.Js * x.Om[i]);
        gu = gu + g * u[i];
    }
    V3 tau = -gu - cross(x.w, mv9(P.I, x.w) + hw);
    V3 wd = mv9(P.Dinv, tau);
    k.w = wd;
    double s2 = dot(x.s, x.s), sw = dot(x.s, x.w);
    k.s = (x.w * (1. - s2) + cross(x.s, x.w) * 2. + x.s * (2. * sw)) * 0.25;
#pragma unroll
    for (int i = 0; i < ON_NRW; i++) k.Om[i] = u[i] * P.invJs - dot(arr(P.gs[i]), wd);
}
ON_HD Rot axpy(const Rot &x, double a, const Rot &k)
{
    Rot o;
    o.s = x.s + k.s * a; o.w = x.w + k.w * a;
#pragma unroll
    for (int i = 0; i < ON_NRW; i++) o.Om[i] = x.Om[i] + k.Om[i] * a;
    return o;
}
ON_HD Rot rk4_rotation(const OpNavParams &P, const Rot &x0, const double (&u)[ON_NRW], double h)
{ // the stage loop stays rolled so that the equations of motion exist once in the instruction stream
    Rot k, xo = x0, x = x0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int st = 0; st < 4; st++) {
        eom_rotation(P, x, u, k);
        const double cw = (st == 0 || st == 3) ? h / 6.0 : h / 3.0;
        const double cn = st == 2 ? h : 0.5 * h;
        xo = axpy(xo, cw, k);
        x = axpy(x0, cn, k);
    }
    return xo;
}

// One state of the bounded random walk of [BSK: utilities/gauss_markov.cpp] computeNextState (identity propagation):
//   x += P * (n + push),  push = e * copysign(e, -x),  e = 1/exp(b^3),  b = max((2 bound - s)/s, 1e-10 bound),
//   s = |x| if |x| > 1e-10 bound else bound.
// For b >= 7.2 the push underflows to zero; for b^3 < 1e-17 it is exactly +-1 (the attitude states: bound 1e-18 deg).
#ifdef __CUDA_ARCH__
// exp(-y), 0 <= y < 400: reduction by ln 2 (two-term Cody-Waite), degree-13 Taylor polynomial on |r| <= ln2/2 (truncation
// 4e-18), exponent patched in; no special operands on this path.
__device__ __forceinline__ double exp_neg(double y)
{
    const double t = fma(y, -1.4426950408889634, 6755399441055744.0);
    const int kk = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, -y);
    r = fma(kd, -1.90821492927058770002e-10, r);
    // degree-13 Taylor polynomial in Estrin form (depth 5 instead of 14)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = r + 1.0, a1 = fma(r, ON_K[29], ON_K[28]), a2 = fma(r, ON_K[31], ON_K[30]), a3 = fma(r, ON_K[33], ON_K[32]);
    const double a4 = fma(r, ON_K[35], ON_K[34]), a5 = fma(r, ON_K[37], ON_K[36]), a6 = fma(r, ON_K[39], ON_K[38]);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double p = fma(fma(a6, r4, b2), r8, fma(b1, r4, b0));
    return __hiloint2double(__double2hiint(p) + (kk << 20), __double2loint(p));
}
#endif
ON_HD double gm_step(double xs, double bound, double pm, double rn)
{
    const double ax = fabs(xs);
    const double sc = ax > bound * 1E-10 ? ax : bound;
#ifdef __CUDA_ARCH__
    double bc = fmad(bound * 2.0, frcp(sc), -1.0);
#else
    double bc = (bound * 2.0 - sc) / sc;
#endif
    bc = bc > bound * 1E-10 ? bc : bound * 1E-10;
    if (bc < 7.2) {
        const double b3 = bc * bc * bc;
#ifdef __CUDA_ARCH__
        const double ex = b3 < 1e-17 ? 1.0 : exp_neg(b3);
#else
        const double ex = b3 < 1e-17 ? 1.0 : 1.0 / exp(b3);
#endif
        rn += ex * copysign(ex, -xs);
    }
    return xs + pm * rn;
}

// ------------------------------------------------------------------------------------------------
// eclipse (conical model, one planet at the origin), coarse sun sensors, cssWlsEst, sunSafePoint
// ------------------------------------------------------------------------------------------------
// Same geometry as leo::penumbra_fraction / leo::eclipse_core (tangent cones of two spheres; the disk-overlap formula of
// eclipse.cpp computePercentShadow inside the guard band around the cone surfaces), for the planet of this scenario.
ON_HD_NOINLINE double penumbra_fraction_on(const OpNavParams &P, double ir, double id, double rdh)
{
    const double PI = 3.14159265358979323846;
    const double ta = P.R_sun * id, tb = P.R_planet * ir;             // sin a, sin b
    const double cc = -rdh * ir * id;                                  // cos c
    const double sc2 = 1. - cc * cc, cb2 = 1. - tb * tb;
    const double sd = (sc2 > 0. && cb2 > 0.) ? sqrt(sc2) * sqrt(cb2) - cc * tb : 2.0;   // sin(c - b)
    if (!(ta <= 0.05 && fabs(sd) <= 0.1 && tb >= 20. * ta && tb < 1.))
        return percent_shadow_general(clamp_asin(ta), clamp_asin(tb), clamp_acos(cc));
    const double a = asin_small(ta), d = asin_small(sd), b = asin(tb);
    if (d < -a) return 0.0;
    if (!(d < a)) return 1.0;
    const double c = b + d;
    const double x = (a * a + d * (2. * b + d)) / (2. * c);
    double y2 = a * a - x * x;
    if (y2 < 0.) y2 = 0.;
    const double y = sqrt(y2);
    const double u = y / b, u2 = u * u;
    double seg = fmad(u2, 5. / 72., 3. / 28.);
    seg = fmad(seg, u2, 1. / 5.); seg = fmad(seg, u2, 2. / 3.);
    const double area = a * a * clamp_acos(x / a) - x * y + (b * b) * (u * u2) * seg;
    return 1. - area / (PI * a * a);
}
// Eclipse::UpdateState for Mars at the origin.  The cone constants follow from |sun| algebraically:
// sin f = (R_sun +- R_p)/|sun|, R_p/sin f = R_p |sun| / (R_sun +- R_p), tan f = sin f / sqrt(1 - sin^2 f).
// Full sun / umbra are decided on squared cone radii; only the band around the cone surfaces evaluates the disk overlap.
ON_HD double eclipse_mars(const OpNavParams &P, V3 sun, V3 r)
{
    const double hp2 = dot(sun, sun), inv_hp = rsq(hp2), hp = hp2 * inv_hp;
    const double Rs = P.R_sun, Rp = P.R_planet;
    const double s1 = (Rs + Rp) * inv_hp, s2f = (Rs - Rp) * inv_hp;
    const double c1off = Rp * hp * P.inv_RsPlusRp, c2off = Rp * hp * P.inv_RsMinusRp;
    const double tan1 = s1 * rsq(1. - s1 * s1), tan2 = s2f * rsq(1. - s2f * s2f);
    const V3 r_HB = sun - r;
    const double s2 = dot(r, r), hb2 = dot(r_HB, r_HB);
    const double s0 = -dot(r, sun) * inv_hp;
    const double c1 = s0 + c1off, c2 = s0 - c2off;
    const double l2sq = s2 - s0 * s0;
    const double l1 = c1 * tan1, l2 = c2 * tan2;
    const double p2 = l1 * l1, u2 = l2 * l2;
    const bool lit = (hb2 < hp2) || (l2sq > p2 * (1. + ECL_BAND) && l2sq > u2 * (1. + ECL_BAND));
    const bool dark = l2sq < u2 * (1. - ECL_BAND) && c2 < 0. && Rs > Rp;
    double f = lit ? 1.0 : 0.0;
    if (!lit && !dark) f = (l2sq < u2 || l2sq < p2) ? penumbra_fraction_on(P, rsq(s2), rsq(hb2), dot(r, r_HB)) : 1.0;
    return f;
}
ON_HD V3 mrp_add(V3 q1, V3 q2)
{ // RigidBodyKinematics addMRP: [FN(out)] = [FB(q2)][BN(q1)], shadow-set guard, inner-set map
    V3 s1 = q1;
    double det = 1. + dot(s1, s1) * dot(q2, q2) - 2. * dot(s1, q2);
    if (fabs(det) < 0.1) {
        s1 = s1 * (-1.0 / dot(s1, s1));
        det = 1. + dot(s1, s1) * dot(q2, q2) - 2. * dot(s1, q2);
    }
    V3 res = s1 * (1. - dot(q2, q2)) + q2 * (1. - dot(s1, s1)) + cross(s1, q2) * 2.;
    return mrp_inner(res * frcp(det));
}
// CSS (prio 299) + cssWlsEst: returns the estimated sun heading (zero when no sensor sees the Sun)
ON_HD_NOINLINE V3 css_wls(const OpNavParams &P, V3 sHat_B, double shadow)
{
    double HtH[6] = {0, 0, 0, 0, 0, 0}, Hty[3] = {0, 0, 0};
    int n = 0;
    V3 H0 = mk(0, 0, 0), H1 = mk(0, 0, 0);
    double y0 = 0, y1 = 0;
    for (int i = 0; i < ON_NCSS; i++) {
        V3 nh = arr(P.cssN[i]);
        double d = dot(nh, sHat_B);
        double y = (d >= P.css_cos_fov ? d : 0.0) * shadow * P.css_scale;
        if (y > 0.0) {
            if (n == 0) { H0 = nh; y0 = y; } else if (n == 1) { H1 = nh; y1 = y; }
            n++;
            HtH[0] += nh.x * nh.x; HtH[1] += nh.x * nh.y; HtH[2] += nh.x * nh.z;
            HtH[3] += nh.y * nh.y; HtH[4] += nh.y * nh.z; HtH[5] += nh.z * nh.z;
            Hty[0] += nh.x * y; Hty[1] += nh.y * y; Hty[2] += nh.z * y;
        }
    }
    V3 d = mk(0, 0, 0);
    if (n >= 3) {
        double m00 = HtH[0], m01 = HtH[1], m02 = HtH[2], m11 = HtH[3], m12 = HtH[4], m22 = HtH[5];
        double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
        double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
        double det = m00 * c00 + m01 * c01 + m02 * c02, id = frcp(det);
        d = mk((c00 * Hty[0] + c01 * Hty[1] + c02 * Hty[2]) * id, (c01 * Hty[0] + c11 * Hty[1] + c12 * Hty[2]) * id,
               (c02 * Hty[0] + c12 * Hty[1] + c22 * Hty[2]) * id);
    } else if (n == 2) {
        double a = dot(H0, H0), b = dot(H0, H1), c = dot(H1, H1), idet = frcp(a * c - b * b);
        double l0 = (c * y0 - b * y1) * idet, l1 = (a * y1 - b * y0) * idet;
        d = H0 * l0 + H1 * l1;
    } else if (n == 1) {
        d = H0 * (y0 * frcp(dot(H0, H0)));
    }
    return n > 0 ? unit_or_zero(d) : mk(0, 0, 0);
}
ON_HD_NOINLINE AttGuid sun_safe_point(V3 sHat, V3 omega_BN_B)
{ // sunSafePoint with sHatBdyCmd = (0,0,1), minUnitMag = smallAngle = sunAxisSpinRate = 0 (ONF:290-295)
    AttGuid g;
    g.omega_RN_B = mk(0, 0, 0); g.domega_RN_B = mk(0, 0, 0); g.sigma_BR = mk(0, 0, 0);
    double sNorm = norm(sHat);
    if (sNorm > 0.0) {
        double ct = sHat.z / sNorm;
        ct = fabs(ct) > 1.0 ? ct / fabs(ct) : ct;
        double err = acos(ct);
        V3 e_hat = unit_or_zero(cross(sHat, mk(0, 0, 1)));
        g.sigma_BR = mrp_inner(e_hat * tan(err * 0.25));
    }
    g.omega_BR_B = omega_BN_B;
    return g;
}

// ------------------------------------------------------------------------------------------------
// relative-OD square-root UKF (see the header comment).  S: lower triangle, row-major, S(i,j) = S[i(i+1)/2+j].
// ------------------------------------------------------------------------------------------------
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
// compile-time loop: f(std::integral_constant<int, I>) for I = B .. E-1 (every index is a constant when the body is generated)
template <int B, int E, class F>
ON_HD void static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
ON_HD void two_body_rk4(double (&x)[6], double mu, double dt)
{
    double k[6], s[6], acc[6];
#pragma unroll
    for (int st = 0; st < 4; st++) {
        const double cin = st == 0 ? 0.0 : (st == 3 ? dt : 0.5 * dt);
        const double cw = (st == 0 || st == 3) ? dt / 6.0 : dt / 3.0;
#pragma unroll
        for (int i = 0; i < 6; i++) s[i] = st == 0 ? x[i] : x[i] + cin * k[i];
        double r2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2], ir = rsq(r2), g = -mu * (ir * ir * ir);
        k[0] = s[3]; k[1] = s[4]; k[2] = s[5]; k[3] = g * s[0]; k[4] = g * s[1]; k[5] = g * s[2];
#pragma unroll
        for (int i = 0; i < 6; i++) acc[i] = st == 0 ? cw * k[i] : acc[i] + cw * k[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] += acc[i];
}
// G independent two-body RK4 steps in lockstep: every operation is emitted G times in a row, so the in-order issue of a warp
// always has G independent FP64 chains to fill the 8-cycle DFMA latency (tests/...: same arithmetic per point as two_body_rk4)
template <int G>
ON_HD void two_body_rk4_group(double (&x)[G][6], double mu, double dt)
{
    double k[G][6], s[G][6], acc[G][6];
#pragma unroll
    for (int st = 0; st < 4; st++) {
        const double cin = st == 0 ? 0.0 : (st == 3 ? dt : 0.5 * dt);
        const double cw = (st == 0 || st == 3) ? dt / 6.0 : dt / 3.0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int g = 0; g < G; g++) s[g][i] = st == 0 ? x[g][i] : x[g][i] + cin * k[g][i];
        double r2[G], ir[G], gg[G];
#pragma unroll
        for (int g = 0; g < G; g++) r2[g] = s[g][0] * s[g][0] + s[g][1] * s[g][1] + s[g][2] * s[g][2];
#pragma unroll
        for (int g = 0; g < G; g++) ir[g] = rsq(r2[g]);
#pragma unroll
        for (int g = 0; g < G; g++) gg[g] = -mu * (ir[g] * ir[g] * ir[g]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int g = 0; g < G; g++) { k[g][i] = s[g][3 + i]; k[g][3 + i] = gg[g] * s[g][i]; }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int g = 0; g < G; g++) acc[g][i] = st == 0 ? cw * k[g][i] : acc[g][i] + cw * k[g][i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int g = 0; g < G; g++) x[g][i] += acc[g][i];
}
// L L^T -= x x^T (hyperbolic sweep); returns false when the result is not positive definite
ON_HD bool chol_downdate6(double (&L)[21], double (&x)[6])
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double Lkk = L[TRI(k, k)], t = Lkk * Lkk - x[k] * x[k];
        if (!(t > 0.0)) { ok = false; t = 1.0; }
        double ir = rsq(t), c = Lkk * ir, s = x[k] * ir;
        L[TRI(k, k)] = t * ir;
#pragma unroll
        for (int j = k + 1; j < 6; j++) {
            double Ljk = L[TRI(j, k)];
            L[TRI(j, k)] = c * Ljk - s * x[j];
            x[j] = c * x[j] - s * Ljk;
        }
    }
    return ok;
}
// Filter state as the kernel holds it during a launch: estimate, square-root covariance packed by columns (column c holds
// rows c .. 5 at C[SC_COL(c) ..]: 21 words; a sigma-point column is read with one run-time base and immediate row offsets),
// mean shift of the last time update.  `Cold` is the cold per-env data of the dynamics side (four Sun nodes of the interval,
// Sun-heading / eclipse latches, nav Sun heading).  The per-thread scratch (Ukf + Cold + Walk = 67 doubles, an odd stride)
// sits in shared memory free of bank conflicts and small enough for three 128-thread blocks per SM; the cold fields are read through volatile accesses so that they are re-loaded at their (rare) points of use
// instead of being carried in registers through the tick loop.
struct Ukf { double x[6]; double C[21]; double m[6]; };                  // 33 doubles
struct Cold { double sun[12]; double cold[7]; };                         // 19 doubles: Sun nodes, cold per-env latches
// simple_nav's 15 walk states: parked in the per-thread shared scratch as well.  The walk advances in a ROLLED loop (four
// states per Philox block; the noise transform exists once in the instruction stream), i.e. the states are indexed at run
// time -- in a per-thread array that means local memory (the kernel's former 752-byte stack: ~140 LDL / STL per tick and
// 2.5x the algorithmic DRAM traffic from its write-backs); in shared memory it is an LDS / STS with an immediate stride.
struct Walk { double e[15]; };
#define SC_COL(c) (6 * (c) - (c) * ((c) - 1) / 2)
#define SC(r, c) C[SC_COL(c) + (r) - (c)]          /* r >= c */
// relODuKFTimeUpdate over dt.  The twelve deviations are accumulated into the 21 independent entries of the Gram
// matrix (no serial dependence between sigma points; propagating the +/- pair of a column side by side was measured and
// is slower: 81 live doubles instead of 57), then one
// 6 x 6 Cholesky factorisation gives the new square-root factor.  FP64 Cholesky of this covariance loses
// eps * cond(scaled P) ~ 1e-12, the same as the differencing of the sigma points themselves.
// Returns false (filter left untouched) when the covariance is not positive definite.
ON_HD_NOINLINE bool ukf_time_update(const OpNavParams &P, Ukf &f, double dt)
{
    double Y0[6], A[21], ms[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { Y0[i] = f.x[i]; ms[i] = 0.0; }
    two_body_rk4(Y0, P.mu_fsw, dt);
#pragma unroll
    for (int i = 0; i < 21; i++) A[i] = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
#ifndef ON_UKF_GROUP
#define ON_UKF_GROUP 4
#endif
    for (int idx = 0; idx < 12; idx += ON_UKF_GROUP) {
        double Y[ON_UKF_GROUP][6];
#pragma unroll
        for (int g = 0; g < ON_UKF_GROUP; g++) {
            const int i = (idx + g) >> 1;
            const double gam = ((idx + g) & 1) ? -P.ukf_gamma : P.ukf_gamma;
#pragma unroll
            for (int r = 0; r < 6; r++) {          // column i of the factor: rows r < i are zero (and not stored)
                const double cr = f.C[SC_COL(i) - i + r];
                Y[g][r] = fmad(gam, r >= i ? cr : 0.0, f.x[r]);
            }
        }
        two_body_rk4_group<ON_UKF_GROUP>(Y, P.mu_fsw, dt);
#pragma unroll
        for (int g = 0; g < ON_UKF_GROUP; g++)
#pragma unroll
            for (int r = 0; r < 6; r++) { Y[g][r] -= Y0[r]; ms[r] += Y[g][r]; }
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b <= a; b++)
#pragma unroll
                for (int g = 0; g < ON_UKF_GROUP; g++) A[TRI(a, b)] = fmad(Y[g][a], Y[g][b], A[TRI(a, b)]);
    }
    double m[6];
#pragma unroll
    for (int r = 0; r < 6; r++) m[r] = P.ukf_w * ms[r];
    const double qp = P.ukf_sq_pos * (dt * dt / 2), qv = P.ukf_sq_vel * dt;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) {
            double v = P.ukf_w * A[TRI(a, b)] + P.ukf_cm * (m[a] * m[b]);
            if (a == b) v += a < 3 ? qp * qp : qv * qv;
            A[TRI(a, b)] = v;
        }
    // 6 x 6 Cholesky, in place on the 21 accumulators.  The triangular loops are unrolled by template recursion, not by
    // `#pragma unroll`: with loop bounds that depend on an outer loop variable the optimiser unrolls too late to promote the
    // arrays to registers and A / L end up in local memory (36 LDL / STL per tick, the kernel's former "stack").
    bool ok = true;
    static_for<0, 6>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        double t = A[TRI(j, j)];
        static_for<0, j>([&](auto kc) { constexpr int k = decltype(kc)::value; t -= A[TRI(j, k)] * A[TRI(j, k)]; });
        if (!(t > 0.0)) { ok = false; t = 1.0; }
        const double ir = rsq(t);
        A[TRI(j, j)] = t * ir;
        static_for<j + 1, 6>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            double v = A[TRI(i, j)];
            static_for<0, j>([&](auto kc) { constexpr int k = decltype(kc)::value; v -= A[TRI(i, k)] * A[TRI(j, k)]; });
            A[TRI(i, j)] = v * ir;
        });
    });
    if (!ok) return false;
#pragma unroll
    for (int c = 0; c < 6; c++)
#pragma unroll
        for (int r = c; r < 6; r++) f.SC(r, c) = A[TRI(r, c)];
#pragma unroll
    for (int i = 0; i < 6; i++) { f.x[i] = Y0[i]; f.m[i] = m[i]; }
    return true;
}
// relODuKFMeasUpdate for y = r_BN_N with noise covariance R (symmetric, m^2, already scaled by noiseSF);
// `dt` is the span of the time update that has just run (its process noise is not part of Pxy / Pyy)
ON_HD_NOINLINE bool ukf_meas_update(const OpNavParams &P, Ukf &f, double dt, const double (&obs)[3], const double (&R)[6])
{
    const double qp = P.ukf_sq_pos * (dt * dt / 2), qp2 = qp * qp;
    double Pxy[6][3];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 3; b++) {
            double s = 0.0;
            const int kmax = a < b ? a : b;
            for (int k = 0; k <= kmax; k++) s += f.SC(a, k) * f.SC(b, k);
            Pxy[a][b] = (a == b) ? s - qp2 : s;
        }
    // Sy = chol(Pyy), Pyy = Pxy[0:3][0:3] + R
    double A00 = Pxy[0][0] + R[0], A10 = Pxy[1][0] + R[1], A20 = Pxy[2][0] + R[2], A11 = Pxy[1][1] + R[3],
           A21 = Pxy[2][1] + R[4], A22 = Pxy[2][2] + R[5];
    if (!(A00 > 0.0)) return false;
    double l00 = sqrt(A00), l10 = A10 / l00, l20 = A20 / l00;
    double t11 = A11 - l10 * l10;
    if (!(t11 > 0.0)) return false;
    double l11 = sqrt(t11), l21 = (A21 - l20 * l10) / l11;
    double t22 = A22 - l20 * l20 - l21 * l21;
    if (!(t22 > 0.0)) return false;
    double l22 = sqrt(t22);
    double innov[3] = {obs[0] - (f.x[0] + f.m[0]), obs[1] - (f.x[1] + f.m[1]), obs[2] - (f.x[2] + f.m[2])};
    double xn[6], U[3][6];
    for (int a = 0; a < 6; a++) {
        // K[a] = (Sy Sy^T)^-1 Pxy[a]: forward then backward substitution
        double t0 = Pxy[a][0] / l00, t1 = (Pxy[a][1] - l10 * t0) / l11, t2 = (Pxy[a][2] - l20 * t0 - l21 * t1) / l22;
        double k2 = t2 / l22, k1 = (t1 - l21 * k2) / l11, k0 = (t0 - l10 * k1 - l20 * k2) / l00;
        xn[a] = f.x[a] + k0 * innov[0] + k1 * innov[1] + k2 * innov[2];
        U[0][a] = k0 * l00 + k1 * l10 + k2 * l20; U[1][a] = k1 * l11 + k2 * l21; U[2][a] = k2 * l22;   // U = K Sy
    }
    double L[21];
    for (int r = 0; r < 6; r++)
        for (int c = 0; c <= r; c++) L[TRI(r, c)] = f.SC(r, c);
    bool ok = true;
    for (int c = 0; c < 3; c++) {
        double x[6];
        for (int a = 0; a < 6; a++) x[a] = U[c][a];
        ok = chol_downdate6(L, x) && ok;
    }
    if (!ok) return false;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c <= r; c++) f.SC(r, c) = L[TRI(r, c)];
    for (int a = 0; a < 6; a++) f.x[a] = xn[a];
    return true;
}

// synthetic camera + circle finder: pinhole image of the Mars disc from the TRUE state (what the renderer shows);
// returns validity (disc centre inside the frame, radius >= houghMinRadius)
ON_HD bool project_circle(const OpNavParams &P, V3 r_C, double (&c)[3])
{
    double d2 = dot(r_C, r_C), Rm = P.planet_radius_km * 1000.0;
    c[0] = c[1] = c[2] = 0.0;
    if (!(r_C.z > 0.0) || !(d2 > Rm * Rm)) return false;
    c[0] = (r_C.x / r_C.z + P.cam_half) / P.cam_X;
    c[1] = (r_C.y / r_C.z + P.cam_half) / P.cam_X;
    c[2] = Rm / sqrt(d2 - Rm * Rm) / P.cam_X;
    return !(c[0] < 0.0 || c[0] >= P.cam_res || c[1] < 0.0 || c[1] >= P.cam_res || c[2] < P.hough_min_radius);
}
// pixelLineConverter (planetTarget = Mars): circle -> r_BN_N [m] and covariance [m^2] (xx, yx, zx, yy, zy, zz) * noiseSF
ON_HD_NOINLINE void pixel_line(const OpNavParams &P, const double (&c)[3], V3 sigma_nav, double (&r_BN_N)[3], double (&R)[6])
{
    const double X = P.cam_X;
    V3 rt = mk(X * c[0] - P.cam_half, X * c[1] - P.cam_half, 1.0);
    V3 rh_C = unit_or_zero(rt) * (-1.0);
    MrpRot CN = mrp_rot(sigma_nav);
    V3 rh_N = rot_NB(CN, sigma_nav, rh_C);
    const double z = X * c[2], q = sqrt(1.0 + z * z);
    const double denom = z / q;                         // sin(atan(z))
    const double rNorm = P.planet_radius_km / denom;    // km
    r_BN_N[0] = rNorm * 1E3 * rh_N.x; r_BN_N[1] = rNorm * 1E3 * rh_N.y; r_BN_N[2] = rNorm * 1E3 * rh_N.z;
    const double x_map = rNorm * X, rho_map = P.planet_radius_km * (X / q - q / (X * c[2] * c[2]));
    const double dxy = x_map * x_map * P.circle_unc, dz = rho_map * rho_map * P.circle_unc;
    // covar_N = [NC] diag(dxy, dxy, dz) [NC]^T = dxy I + (dz - dxy) n n^T with n = [NC] e_z
    V3 nz = rot_NB(CN, sigma_nav, mk(0, 0, 1));
    const double sc = 1E6 * P.ukf_noiseSF, a = dxy * sc, b = (dz - dxy) * sc;
    R[0] = a + b * nz.x * nz.x; R[1] = b * nz.y * nz.x; R[2] = b * nz.z * nz.x;
    R[3] = a + b * nz.y * nz.y; R[4] = b * nz.z * nz.y; R[5] = a + b * nz.z * nz.z;
}

// ------------------------------------------------------------------------------------------------
// initial conditions (ONS:163-202) and reset (scenario_OpNav.__init__, ONE:153-168)
// ------------------------------------------------------------------------------------------------
ON_HD_NOINLINE void sample_ic(const OpNavParams &P, int64_t global_env, int64_t episode, double (&ic)[OPNAV_IC_DIM])
{
    IcRng g; g.seed = P.seed; g.env = (uint64_t)global_env; g.episode = (uint32_t)episode; g.block = 0; g.have = 0;
    const double PI = 3.14159265358979323846, D2R = PI / 180.0;
    if (P.sample_orbit) { // the commented-out draws of ONS:166-171
        double a = g.uniform(17000 * 1E3, 22000 * 1E3), ecc = g.uniform(0, 0.6), inc = g.uniform(-20 * D2R, 20 * D2R);
        double Om = g.uniform(0 * D2R, 360 * D2R), om = g.uniform(0 * D2R, 360 * D2R), f = g.uniform(0 * D2R, 360 * D2R);
        double mu = P.mu_dyn, p = a * (1.0 - ecc * ecc), r = p / (1.0 + ecc * cos(f)), th = om + f, h = sqrt(mu * p);
        ic[0] = r * (cos(th) * cos(Om) - cos(inc) * sin(th) * sin(Om));
        ic[1] = r * (cos(th) * sin(Om) + cos(inc) * sin(th) * cos(Om));
        ic[2] = r * (sin(th) * sin(inc));
        ic[3] = -mu / h * (cos(Om) * (ecc * sin(om) + sin(th)) + cos(inc) * (ecc * cos(om) + cos(th)) * sin(Om));
        ic[4] = -mu / h * (sin(Om) * (ecc * sin(om) + sin(th)) - cos(inc) * (ecc * cos(om) + cos(th)) * cos(Om));
        ic[5] = mu / h * (ecc * cos(om) + cos(th)) * sin(inc);
    } else {
        for (int k = 0; k < 3; k++) { ic[k] = P.rN0[k]; ic[3 + k] = P.vN0[k]; }
    }
    for (int k = 0; k < 3; k++) ic[6 + k] = g.uniform(100000, -100000);     // ONS:187
    for (int k = 0; k < 3; k++) ic[9 + k] = g.uniform(1000, -1000);         // ONS:188
}
ON_HD void opnav_reset_env(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e,
                           const double (&ic)[OPNAV_IC_DIM], double *obs /* 4, may be null */)
{
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
    for (int f = 0; f < OPNAV_ND; f++) SD(f) = 0.0;
    int64_t episode = SI(OI_EPISODE);
    for (int f = 0; f < OPNAV_NI; f++) SI(f) = 0;
    SI(OI_EPISODE) = episode;
    for (int k = 0; k < 6; k++) SD(OF_R + k) = ic[k];                        // sigma = omega = Omega = 0 (ONS:189-194)
    for (int k = 0; k < 6; k++) SD(OF_FSTATE + k) = ic[k] + ic[6 + k];       // stateInit (ONS:189)
    for (int k = 0; k < 3; k++) { SD(OF_FS + TRI(k, k)) = P.ukf_P0_pos; SD(OF_FS + TRI(k + 3, k + 3)) = P.ukf_P0_vel; }
    SD(OF_SHADOW) = 1.0;
    SI(OI_TICK) = -1;
    SI(OI_CAMERA) = 1;                                                       // OND:130
    SI(OI_FIRST) = 1;
    if (obs) for (int k = 0; k < 4; k++) obs[k] = 0.0;                       // ONS:152
#undef SD
#undef SI
}

// ------------------------------------------------------------------------------------------------
// one decision interval of one environment
// ------------------------------------------------------------------------------------------------
struct StepOut { double ob[4]; double debug[12]; double reward; int done; int reason; };

// A tick has three dependency chains of similar length that only meet through small messages:
//   NoiseRole   the bounded random walk of simple_nav's 15 error states (depends on nothing but its own stream);
//   DynRole     wheel latch, CSS, eclipse, truth RK4, nav message, camera, guidance, control (and the circle -> pixelLine
//               measurement when a frame is due): consumes the error states, produces the measurement;
//   FilterRole  relativeODuKF: one time update per tick, a measurement update when the dynamics side produced one.
// opnav_step_env() runs them in sequence in one thread (host-compiled core); the warp-specialised kernel of opnav.cu gives
// them to three warps of a block, each one tick behind the previous, with the messages in shared mailboxes.
struct Meas { bool valid; double obs[3]; double R[6]; };
// Hand-over between the two passes of a decision interval: slot-major [slot][ON_MEAS_W][env] doubles.  Slot 0 is a header
// (number of measurements, nav Sun heading of the last tick, episode flags, tick range), slots 1.. hold the measurements
// (tick, obs[3], R[6]) of the interval's valid camera frames.
#define ON_MEAS_W 10
enum OnHdr : int { OH_NM = 0, OH_SUN = 1, OH_OVER = 4, OH_REASON = 5, OH_KFIRST = 6, OH_KLAST = 7 };
struct MeasBuf { double *p; int64_t stride; };

// simple_nav's Gauss-Markov error states: a bounded random walk driven by the per-env Philox stream, independent of the
// dynamics (SimpleNav::computeErrors, OND:236-258).  tick(k) advances the 15 states to tick k.
struct NoiseRole {
    double *nerr;                 // Walk::e of this env (shared scratch on the device)
    int64_t genv, episode, k_first, k_last;
    ON_HD void load(const OpNavParams &P, const double *S, const int64_t *I, int64_t stride, int64_t e, Walk &w)
    {
        nerr = w.e;
        for (int i = 0; i < 15; i++) nerr[i] = S[(int64_t)(OF_NAVERR + i) * stride + e];
        genv = P.first_env_index + e; episode = I[(int64_t)OI_EPISODE * stride + e];
        const int64_t tick0 = I[(int64_t)OI_TICK * stride + e];
        k_first = tick0 + 1; k_last = (tick0 < 0 ? 0 : tick0) + P.ticks_per_step;
    }
    ON_HD void tick(const OpNavParams &P, int64_t k)
    {
        if (!P.nav_noise) return;
        const double ndt = k > 0 ? P.dt : 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) nerr[i] += ndt * nerr[3 + i];
#ifndef ON_NZ_UNROLL
#define ON_NZ_UNROLL 1
#endif
#if defined(__CUDA_ARCH__)
#if ON_NZ_UNROLL == 1
#pragma unroll 1
#elif ON_NZ_UNROLL == 2
#pragma unroll 2
#else
#pragma unroll
#endif
#endif
        for (int b = 0; b < 4; b++) {                                 // four normals per Philox block, 15 walk states
            double n4[4];
            normals4_inl(P, genv, episode, (uint32_t)k, 1u, (uint32_t)b, n4);
#pragma unroll
            for (int j = 0; j < 4; j++) {                             // four independent walk states side by side
                const int i = 4 * b + j;
                if (i < 15) nerr[i] = gm_step(nerr[i], P.navBound[i], P.navP[i], n4[j]);
            }
        }
    }
    ON_HD void finish(double *S, int64_t stride, int64_t e) const
    {
        for (int i = 0; i < 15; i++) S[(int64_t)(OF_NAVERR + i) * stride + e] = nerr[i];
    }
};

struct DynRole {
    Truth x;
    double rwcmd[ON_NRW];
    int mode, camera, sunpt_w, over, reason, modeCounter;
    int64_t n_img, n_switch, curr_step, k_first, k_last, genv, episode;
    double sun_t0, sun_T;

    ON_HD void load(const OpNavParams &P, const double *S, const int64_t *I, int64_t stride, int64_t e, int action, Cold &c)
    {
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
        genv = P.first_env_index + e; episode = SI(OI_EPISODE);
        // ---- opNavEnv.step prologue (ONE:94-95) and run_sim mode switching (ONS:237-254) ----
        over = (int)SI(OI_OVER); reason = 0;
        curr_step = SI(OI_STEP);
        if (curr_step >= P.max_length) { over = 1; reason |= 1; }
        mode = (int)SI(OI_MODE); camera = (int)SI(OI_CAMERA);
        modeCounter = (int)SI(OI_MODECNT) + 1;
        if (action == 0) { mode = 0; if (P.camera_reenable) camera = 1; }
        else if (action == 1) { mode = 1; camera = 0; }
        if (SI(OI_FIRST)) mode = 0;      // pending 'OpNavOD' event fires at the first ExecuteSimulation (ONS:157, ONF:219-224)
        x.r = mk(SD(OF_R), SD(OF_R + 1), SD(OF_R + 2)); x.v = mk(SD(OF_V), SD(OF_V + 1), SD(OF_V + 2));
        x.s = mk(SD(OF_SIG), SD(OF_SIG + 1), SD(OF_SIG + 2)); x.w = mk(SD(OF_OMG), SD(OF_OMG + 1), SD(OF_OMG + 2));
        for (int i = 0; i < ON_NRW; i++) { x.Om[i] = SD(OF_WHL + i); rwcmd[i] = SD(OF_RWCMD + i); }
        volatile double *cold = c.cold;        // [0..2] sun_point_data, [3] eclipse message, [4..6] nav Sun heading
        cold[0] = SD(OF_SUNPT); cold[1] = SD(OF_SUNPT + 1); cold[2] = SD(OF_SUNPT + 2);
        sunpt_w = (int)SI(OI_SUNPT_W);
        cold[3] = SD(OF_SHADOW);
        cold[4] = 0.0; cold[5] = 0.0; cold[6] = 1.0;
        n_img = SI(OI_NIMG); n_switch = SI(OI_SWITCH);
        const int64_t tick0 = SI(OI_TICK);
        k_first = tick0 + 1; k_last = (tick0 < 0 ? 0 : tick0) + P.ticks_per_step;   // stop time inclusive
        // ---- Sun over this interval ----
        sun_t0 = (double)(tick0 < 0 ? 0 : tick0) * P.dt; sun_T = (double)P.ticks_per_step * P.dt;
        volatile double *sunn = c.sun;
        for (int j = 0; j < 4; j++) {
            V3 p = sun_from_mars(P, sun_t0 + (double)j * sun_T / 3.0).r;
            sunn[3 * j] = p.x; sunn[3 * j + 1] = p.y; sunn[3 * j + 2] = p.z;
        }
#undef SD
#undef SI
    }

    // one tick of the dynamics process and of the flight software except the filter; `nerr` = simple_nav's error states at
    // this tick (NoiseRole), `m` receives the measurement of this tick
    ON_HD void tick(const OpNavParams &P, int64_t k, Cold &c, const double *nerr, Meas &m)
    {
        volatile double *cold = c.cold;
        volatile double *sunn = c.sun;
        const double t = (double)k * P.dt;
        // ================= DynamicsTask =================
        // ReactionWheelStateEffector (prio 301): latch last pass's motor torques; publish Omega before the integration
        double u[ON_NRW], Om_msg[ON_NRW];
#pragma unroll
        for (int i = 0; i < ON_NRW; i++) {
            double ui = rwcmd[i];
            ui = ui > P.u_max ? P.u_max : (ui < -P.u_max ? -P.u_max : ui);
            if (fabs(x.Om[i]) >= P.Om_max && x.Om[i] * ui >= 0.0) ui = 0.0;
            u[i] = ui; Om_msg[i] = x.Om[i];
        }
        // CSSConstellation (prio 299) then Eclipse (prio 204): both see the spacecraft / Sun messages of the previous
        // tick; the CSS see the eclipse message of the previous tick.  Only the sun-safe pass consumes the CSS.
        V3 css_sun = mk(0, 0, 0);
        if (mode == 1) {
            if (k > 0) {
                MrpRot BN = mrp_rot(x.s);
                css_sun = css_wls(P, rot_BN(BN, x.s, unit_or_zero(sun_at(sunn, sun_t0, sun_T, t - P.dt) - x.r)), cold[3]);
            }
        }
        if (mode == 1 || k == k_last) cold[3] = k > 0 ? eclipse_mars(P, sun_at(sunn, sun_t0, sun_T, t - P.dt), x.r) : 1.0;
        // SpacecraftPlus (prio 201)
        if (k > 0) {
            rk4_translation(P, x.r, x.v, P.dt);
            Rot q;
            q.s = x.s; q.w = x.w;
#pragma unroll
            for (int i = 0; i < ON_NRW; i++) q.Om[i] = x.Om[i];
            q = rk4_rotation(P, q, u, P.dt);
            x.s = q.s; x.w = q.w;
#pragma unroll
            for (int i = 0; i < ON_NRW; i++) x.Om[i] = q.Om[i];
            double s2 = dot(x.s, x.s);
            if (s2 > 1.0) { x.s = x.s * (-1.0 / s2); n_switch++; }      // |sigma| > 1
        }
        // SpiceInterface (prio 200): the Sun message of this tick is evaluated where it is consumed (here at the last tick
        // for the nav Sun heading, at the next tick by the CSS and the eclipse model)
        // SimpleNav (prio 109): error states from the noise role, applied to the truth
        const V3 nav_r = x.r + mk(nerr[0], nerr[1], nerr[2]), nav_v = x.v + mk(nerr[3], nerr[4], nerr[5]);
        const V3 nav_s = P.nav_noise ? mrp_add(x.s, mk(nerr[6], nerr[7], nerr[8])) : mrp_inner(x.s);
        const V3 nav_w = x.w + mk(nerr[9], nerr[10], nerr[11]);
        if (k == k_last) {
            MrpRot BN = mrp_rot(x.s);
            V3 sb = rot_BN(BN, x.s, unit_or_zero(sun_at(sunn, sun_t0, sun_T, t) - x.r));
            V3 se = mk(nerr[12], nerr[13], nerr[14]);
            MrpRot OT = mrp_rot(se);
            V3 ns = rot_BN(OT, se, sb);
            cold[4] = ns.x; cold[5] = ns.y; cold[6] = ns.z;
        }
        // CameraTask (prio 999, every cam_ticks): the frame shows the true state of this tick
        const bool frame = camera && (k % P.cam_ticks == 0);
        if (frame) n_img++;
        // ================= FSW process =================
        AttGuid g;
        if (mode == 0) { // opNavPointTaskCheat: hillPoint + attTrackingError(sigma_R0R)
            AttRef ref = hill_point(nav_r, nav_v, mk(0, 0, 0), mk(0, 0, 0));
            ref.sigma_RN = mrp_add(ref.sigma_RN, arr(P.sigma_RR0));
            g.sigma_BR = mrp_sub(nav_s, ref.sigma_RN);
            MrpRot BN = mrp_rot(nav_s);
            g.omega_RN_B = rot_BN(BN, nav_s, ref.omega_RN_N);
            g.omega_BR_B = nav_w - g.omega_RN_B;
            g.domega_RN_B = rot_BN(BN, nav_s, ref.domega_RN_N);
        } else {         // sunSafePointTask: sunSafePoint (last pass's heading) then cssWlsEst
            g = sun_safe_point(sunpt_w ? mk(cold[0], cold[1], cold[2]) : mk(0, 0, 0), nav_w);
            cold[0] = css_sun.x; cold[1] = css_sun.y; cold[2] = css_sun.z; sunpt_w = 1;
        }
        { // mrpFeedbackRWsTask: MRP_Feedback with wheel momentum, rwMotorTorque
            V3 w = g.omega_BR_B + g.omega_RN_B;
            V3 Lr = g.omega_BR_B * P.Pgain + g.sigma_BR * P.K;
            V3 hI = mv9(P.I, w);
#pragma unroll
            for (int i = 0; i < ON_NRW; i++) {
                V3 gi = arr(P.gs[i]);
                hI = hI + gi * (P.Js * (dot(w, gi) + Om_msg[i]));
            }
            Lr = Lr - cross(g.omega_RN_B, hI);
            Lr = Lr - mv9(P.I, g.domega_RN_B - cross(w, g.omega_RN_B));
            // commanded torque = -Lr; rwMotorTorque maps -(commanded) = Lr
#pragma unroll
            for (int i = 0; i < ON_NRW; i++) rwcmd[i] = dot(arr(P.Umap[i]), Lr);
        }
        // opNavODTask front end: imageProcessing stand-in + pixelLine (the filter itself is the other role)
        m.valid = false;
        if (mode == 0 && frame) {
            double cc[3];
            MrpRot BN = mrp_rot(x.s);
            m.valid = project_circle(P, rot_BN(BN, x.s, -x.r), cc);
            if (m.valid) {
                if (P.pixel_noise_std > 0.0) {
                    double n4[4];
                    normals4(P, genv, episode, (uint32_t)k, 2u, 0u, n4);
                    cc[0] += P.pixel_noise_std * n4[0]; cc[1] += P.pixel_noise_std * n4[1]; cc[2] += P.pixel_noise_std * n4[2];
                }
                pixel_line(P, cc, nav_s, m.obs, m.R);
            }
        }
    }

    // The dynamics side of the persistent state at the end of the first pass; what the end of the interval needs besides it
    // (nav Sun heading, episode flags, the tick range) goes into the header slot of the measurement buffer (opnav_finish_obs).
    ON_HD void finish_state(double *S, int64_t *I, int64_t stride, int64_t e, Cold &c, double *hdr, int64_t hs, int n_m)
    {
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
        volatile double *cold = c.cold;
        SD(OF_R) = x.r.x; SD(OF_R + 1) = x.r.y; SD(OF_R + 2) = x.r.z; SD(OF_V) = x.v.x; SD(OF_V + 1) = x.v.y; SD(OF_V + 2) = x.v.z;
        SD(OF_SIG) = x.s.x; SD(OF_SIG + 1) = x.s.y; SD(OF_SIG + 2) = x.s.z; SD(OF_OMG) = x.w.x; SD(OF_OMG + 1) = x.w.y; SD(OF_OMG + 2) = x.w.z;
        for (int i = 0; i < ON_NRW; i++) { SD(OF_WHL + i) = x.Om[i]; SD(OF_RWCMD + i) = rwcmd[i]; }
        SD(OF_SUNPT) = cold[0]; SD(OF_SUNPT + 1) = cold[1]; SD(OF_SUNPT + 2) = cold[2];
        SD(OF_SHADOW) = cold[3];
        SI(OI_TICK) = k_last; SI(OI_STEP) = curr_step + 1; SI(OI_MODE) = mode; SI(OI_CAMERA) = camera; SI(OI_MODECNT) = modeCounter;
        SI(OI_FIRST) = 0; SI(OI_SWITCH) = n_switch; SI(OI_SUNPT_W) = sunpt_w; SI(OI_NIMG) = n_img;
        hdr[OH_NM * hs] = (double)n_m;
        hdr[(OH_SUN + 0) * hs] = cold[4]; hdr[(OH_SUN + 1) * hs] = cold[5]; hdr[(OH_SUN + 2) * hs] = cold[6];
        hdr[OH_OVER * hs] = (double)over; hdr[OH_REASON * hs] = (double)reason;
        hdr[OH_KFIRST * hs] = (double)k_first; hdr[OH_KLAST * hs] = (double)k_last;
#undef SD
#undef SI
    }
};

struct FilterRole {
    int64_t ftick, n_meas, n_bad, k_first, k_last;

    ON_HD void load(const OpNavParams &P, const double *S, const int64_t *I, int64_t stride, int64_t e, Ukf &f, int64_t kf, int64_t kl)
    {
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
        for (int i = 0; i < 6; i++) { f.x[i] = SD(OF_FSTATE + i); f.m[i] = 0.0; }
        for (int r = 0; r < 6; r++)
            for (int c = 0; c <= r; c++) f.SC(r, c) = SD(OF_FS + TRI(r, c));
        ftick = SI(OI_FTICK); n_meas = SI(OI_NMEAS); n_bad = SI(OI_NBAD);
        k_first = kf; k_last = kl;       // from the first pass (which has already advanced OI_TICK)
#undef SD
#undef SI
    }
    // relativeODuKF (opNavODTask: after imageProcessing + pixelLine; sunSafePointTask: time update only)
    ON_HD void tick(const OpNavParams &P, Ukf &f, int64_t k, bool meas, const double (&obs)[3], const double (&R)[6])
    {
        const double fdt = (double)(k - ftick) * P.dt;
        bool tu_ok = true;
        if (meas || k > ftick) {
            tu_ok = ukf_time_update(P, f, fdt);
            if (tu_ok) ftick = k; else n_bad++;              // relODuKFCleanUpdate: the filter keeps its previous state and time
        }
        if (meas && tu_ok) {
            if (ukf_meas_update(P, f, fdt, obs, R)) n_meas++; else n_bad++;
        }
    }
    ON_HD void finish(double *S, int64_t *I, int64_t stride, int64_t e, const Ukf &f, double (&fx)[3], double (&psig)[3])
    {
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
        fx[0] = f.x[0]; fx[1] = f.x[1]; fx[2] = f.x[2];
        psig[0] = sqrt(f.SC(0, 0) * f.SC(0, 0));
        psig[1] = sqrt(f.SC(1, 0) * f.SC(1, 0) + f.SC(1, 1) * f.SC(1, 1));
        psig[2] = sqrt(f.SC(2, 0) * f.SC(2, 0) + f.SC(2, 1) * f.SC(2, 1) + f.SC(2, 2) * f.SC(2, 2));
        for (int i = 0; i < 6; i++) SD(OF_FSTATE + i) = f.x[i];
        for (int r = 0; r < 6; r++)
            for (int c = 0; c <= r; c++) SD(OF_FS + TRI(r, c)) = f.SC(r, c);
        SI(OI_NMEAS) = n_meas; SI(OI_NBAD) = n_bad; SI(OI_FTICK) = ftick;
#undef SD
#undef SI
    }
};

// Both roles in one thread (host-compiled core; single-role kernel).  `f` and `c` are scratch storage for the call.
ON_HD int opnav_meas_slots(const OpNavParams &P) { return P.ticks_per_step / (P.cam_ticks > 0 ? P.cam_ticks : 1) + 3; }   // + header

// One decision interval of one env runs in TWO PASSES over its ticks: first the noise walk and the dynamics / flight-software
// role, which leave the interval's measurements (one per valid camera frame) and a header in `mb`; then the filter, which
// between frames needs nothing from the other two (FilterRole), and the observation / reward / termination of the step.  Same
// arithmetic, same order per role, as a single interleaved loop; but the working sets of the two passes are never live
// together: they are separate kernels (opnav.cu: opnav_pass1_kernel, opnav_pass2_kernel), each with its own register
// allocation and shared-memory scratch, three blocks per SM.  DESIGN.md 6b.
ON_HD void opnav_pass1(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e, int action, Cold &c, Walk &w, MeasBuf mb)
{
    DynRole d;
    NoiseRole nz;
    d.load(P, S, I, stride, e, action, c);
    nz.load(P, S, I, stride, e, w);
    int n_m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int64_t k = d.k_first; k <= d.k_last; k++) {
        Meas m;
        // (ON_EXP_*: measurement switches -- compile one role out to read its cost off the clock; never in the product)
#ifdef ON_EXP_NODYN
        m.valid = false;
#endif
#ifndef ON_EXP_NONOISE
        nz.tick(P, k);
#endif
#ifndef ON_EXP_NODYN
        d.tick(P, k, c, nz.nerr, m);
#endif
        if (m.valid) {
            n_m++;
            double *q = mb.p + (int64_t)n_m * ON_MEAS_W * mb.stride;
            q[0] = (double)k;
            for (int i = 0; i < 3; i++) q[(1 + i) * mb.stride] = m.obs[i];
            for (int i = 0; i < 6; i++) q[(4 + i) * mb.stride] = m.R[i];
        }
    }
    nz.finish(S, stride, e);
    d.finish_state(S, I, stride, e, c, mb.p, mb.stride, n_m);
}

// THREE-KERNEL form of the interval (opnav.cu: opnav_noise_kernel, opnav_dyn_kernel, opnav_pass2_kernel).  The noise walk
// depends on nothing but its own Philox stream, so the first pass itself splits: `opnav_pass0` advances the 15 error states
// through the interval and leaves them, tick by tick, in a slot-major global buffer (15 doubles per env-tick: 360 KB per env
// and interval -- HBM is 180 GB); `opnav_pass1_fed` is the first pass with the walk replaced by that feed, double-buffered in
// the per-thread shared scratch one tick ahead (cp.async on the device: no registers, no exposed load latency).  Same
// arithmetic per role as opnav_pass1: results are bit-identical.  Each kernel gets the register allocation and residency
// of its own working set (the walk: 15 states and a Philox block; the dynamics: truth, flight software, Sun nodes).
ON_HD void opnav_pass0(const OpNavParams &P, double *S, const int64_t *I, int64_t stride, int64_t e, Walk &w, double *nz, int64_t nz_stride)
{
    NoiseRole nzr;
    nzr.load(P, S, I, stride, e, w);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int64_t k = nzr.k_first; k <= nzr.k_last; k++) {
        nzr.tick(P, k);
        double *q = nz + (k - nzr.k_first) * 15 * nz_stride;
#pragma unroll
        for (int i = 0; i < 15; i++) q[(int64_t)i * nz_stride] = nzr.nerr[i];
    }
    nzr.finish(S, stride, e);
}

struct NoiseFeed {
    const double *g;              // this slot's column of the noise buffer: value i of local tick kl at g[(kl * 15 + i) * stride]
    int64_t stride;
    double *buf;                  // [2][15] in the per-thread scratch
    ON_HD void prefetch(int64_t kl) const
    {
        double *dst = buf + (kl & 1) * 15;
        const double *src = g + kl * 15 * stride;
#if defined(__CUDA_ARCH__)
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int i = 0; i < 15; i++)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" : : "r"(d + 8u * i), "l"(src + (int64_t)i * stride) : "memory");
#else
        for (int i = 0; i < 15; i++) dst[i] = src[(int64_t)i * stride];
#endif
    }
    ON_HD const double *get(int64_t kl) const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.wait_all;" : : : "memory");
#endif
        return buf + (kl & 1) * 15;
    }
};

ON_HD void opnav_pass1_fed(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e, int action, Cold &c, NoiseFeed nf, MeasBuf mb)
{
    DynRole d;
    d.load(P, S, I, stride, e, action, c);
    int n_m = 0;
    nf.prefetch(0);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int64_t k = d.k_first; k <= d.k_last; k++) {
        Meas m;
        const int64_t kl = k - d.k_first;
        const double *nerr = nf.get(kl);
        if (k < d.k_last) nf.prefetch(kl + 1);
        d.tick(P, k, c, nerr, m);
        if (m.valid) {
            n_m++;
            double *q = mb.p + (int64_t)n_m * ON_MEAS_W * mb.stride;
            q[0] = (double)k;
            for (int i = 0; i < 3; i++) q[(1 + i) * mb.stride] = m.obs[i];
            for (int i = 0; i < 6; i++) q[(4 + i) * mb.stride] = m.R[i];
        }
    }
    d.finish_state(S, I, stride, e, c, mb.p, mb.stride, n_m);
}

// observation (ONS:263-293) and opNavEnv.step epilogue (ONE:100-125, :139-152); fx = filter position estimate, psig = sqrt of the
// first three covariance diagonal entries; the truth comes back from the state the first pass stored
ON_HD void opnav_finish_obs(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e, int action, const double (&fx)[3],
                            const double (&psig)[3], MeasBuf mb, StepOut &out)
{
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
    Truth x;
    x.r = mk(SD(OF_R), SD(OF_R + 1), SD(OF_R + 2)); x.v = mk(SD(OF_V), SD(OF_V + 1), SD(OF_V + 2));
    x.s = mk(SD(OF_SIG), SD(OF_SIG + 1), SD(OF_SIG + 2));
    int over = (int)mb.p[OH_OVER * mb.stride], reason = (int)mb.p[OH_REASON * mb.stride];
    const int modeCounter = (int)SI(OI_MODECNT);
    const double nr2 = fx[0] * fx[0] + fx[1] * fx[1] + fx[2] * fx[2], inr = 1.0 / sqrt(nr2);
    {
        MrpRot BN = mrp_rot(x.s);
        V3 pos_B = -rot_BN(BN, x.s, mk(fx[0], fx[1], fx[2]) * inr);
        V3 nav_sun_B = mk(mb.p[(OH_SUN + 0) * mb.stride], mb.p[(OH_SUN + 1) * mb.stride], mb.p[(OH_SUN + 2) * mb.stride]);
        V3 sh = nav_sun_B * (1.0 / norm(nav_sun_B));
        out.ob[0] = dot(pos_B, sh);
        out.ob[1] = psig[0] * inr; out.ob[2] = psig[1] * inr; out.ob[3] = psig[2] * inr;
    }
    out.debug[0] = fx[0]; out.debug[1] = fx[1]; out.debug[2] = fx[2];
    out.debug[3] = x.r.x; out.debug[4] = x.r.y; out.debug[5] = x.r.z;
    out.debug[6] = x.v.x; out.debug[7] = x.v.y; out.debug[8] = x.v.z;
    out.debug[9] = x.s.x; out.debug[10] = x.s.y; out.debug[11] = x.s.z;
    double reward = 0.0;
    if (action == 1) {
        V3 real = x.r, nav = (mk(fx[0], fx[1], fx[2]) - real) * (1.0 / norm(real));
        reward = fabs(P.reward_mult / (1.0 + dot(nav, nav)));
    }
    if (modeCounter >= P.numModes) { over = 1; reason |= 2; }
    out.reward = reward; out.done = over; out.reason = reason;
    SD(OF_EPRET) = SD(OF_EPRET) + reward;
    for (int i = 0; i < 4; i++) SD(OF_OBS + i) = out.ob[i];
    for (int i = 0; i < 12; i++) SD(OF_DEBUG + i) = out.debug[i];
    SI(OI_OVER) = over;
#undef SD
#undef SI
}

ON_HD void opnav_pass2(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e, int action, StepOut &out, Ukf &f, MeasBuf mb)
{
    FilterRole fr;
    fr.load(P, S, I, stride, e, f, (int64_t)mb.p[OH_KFIRST * mb.stride], (int64_t)mb.p[OH_KLAST * mb.stride]);
#ifndef ON_EXP_NOFILTER
    const int n_m = (int)mb.p[OH_NM * mb.stride];
    int next = 1;
    double k_next = n_m > 0 ? mb.p[(int64_t)ON_MEAS_W * mb.stride] : -1.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int64_t k = fr.k_first; k <= fr.k_last; k++) {
        double obs[3] = {0., 0., 0.}, R[6] = {0., 0., 0., 0., 0., 0.};
        const bool meas = next <= n_m && k_next == (double)k;
        if (meas) {
            const double *q = mb.p + (int64_t)next * ON_MEAS_W * mb.stride;
            for (int i = 0; i < 3; i++) obs[i] = q[(1 + i) * mb.stride];
            for (int i = 0; i < 6; i++) R[i] = q[(4 + i) * mb.stride];
            next++;
            k_next = next <= n_m ? mb.p[(int64_t)next * ON_MEAS_W * mb.stride] : -1.0;
        }
        fr.tick(P, f, k, meas, obs, R);
    }
#endif
    double fx[3], psig[3];
    fr.finish(S, I, stride, e, f, fx, psig);
    opnav_finish_obs(P, S, I, stride, e, action, fx, psig, mb, out);
}

// both passes in one thread (the host-compiled core of the tests)
ON_HD void opnav_step_env(const OpNavParams &P, double *S, int64_t *I, int64_t stride, int64_t e, int action, StepOut &out,
                          Ukf &f, Cold &c, Walk &w, MeasBuf mb)
{
    opnav_pass1(P, S, I, stride, e, action, c, w, mb);
    opnav_pass2(P, S, I, stride, e, action, out, f, mb);
}

}  // namespace opnav
