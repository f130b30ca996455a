// opnav_host.h -- host-side derivation of OpNavParams from bskenv_opnav_config: the scenario wiring of
//   /root/reference/basilisk_env/simulators/opNavSimulator.py                  (ONS:line)
//   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavDynamics.py  (OND:line)
//   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavFsw.py       (ONF:line)
//   /root/reference/basilisk_env/envs/opNavEnvironment.py                      (ONE:line)
// turned into numbers; Basilisk factory values are cited [BSK].
#pragma once
#include <math.h>
#include <string.h>
#include <string>
#include "../../include/bskenv.h"
#include "leo_host.h"
#include "opnav_params.h"

namespace opnav_host {

static inline void default_config(bskenv_opnav_config *c)
{
    memset(c, 0, sizeof(*c));
    c->abi_version = BSKENV_ABI_VERSION;
    c->dynRate = 1.0; c->fswRate = 1.0; c->step_duration_min = 50.0;       // ONE:86, ONE:31
    c->max_length = 40; c->numModes = 50;                                   // ONE:23, ONS:149
    c->nav_noise = 1; c->camera_reenable = 0; c->sample_orbit = 0; c->auto_reset = 0;
    c->pixel_noise_std = 0.5; c->circle_unc = 0.25;
    c->reward_mult = 1.0;                                                   // ONE:32
    c->noise_seed = 0;
}

static inline void elem2rv(double mu, double a, double e, double i, double Om, double om, double f, double r[3], double v[3])
{ // [BSK: orbitalMotion.elem2rv], non-rectilinear branch
    double p = a * (1.0 - e * e), rr = p / (1.0 + e * cos(f)), th = om + f, h = sqrt(mu * p);
    r[0] = rr * (cos(th) * cos(Om) - cos(i) * sin(th) * sin(Om));
    r[1] = rr * (cos(th) * sin(Om) + cos(i) * sin(th) * cos(Om));
    r[2] = rr * (sin(th) * sin(i));
    v[0] = -mu / h * (cos(Om) * (e * sin(om) + sin(th)) + cos(i) * (e * cos(om) + cos(th)) * sin(Om));
    v[1] = -mu / h * (sin(Om) * (e * sin(om) + sin(th)) - cos(i) * (e * cos(om) + cos(th)) * cos(Om));
    v[2] = mu / h * (e * cos(om) + cos(th)) * sin(i);
}

static inline std::string build_params(const bskenv_opnav_config &c, OpNavParams &p)
{
    memset(&p, 0, sizeof(p));
    if (c.abi_version != BSKENV_ABI_VERSION) return "bskenv_opnav_config.abi_version mismatch";
    if (!(c.dynRate > 0) || !(c.step_duration_min > 0)) return "rates must be positive";
    if (c.fswRate != c.dynRate) return "the opNav env runs dynamics and flight software at the same rate (ONE:86)";
    const double PI = 3.14159265358979323846, D2R = PI / 180.0, RPM = 0.10471975511965977;
    p.dyn_ns = leo_host::sec2nano(c.dynRate); p.dt = c.dynRate;
    const int64_t step_ns = (int64_t)(c.step_duration_min * 60.0 * 1e9 + 0.5);       // mc.min2nano (ONS:257)
    if (step_ns % p.dyn_ns) return "step_duration must be a multiple of dynRate";
    p.ticks_per_step = (int32_t)(step_ns / p.dyn_ns);
    const int64_t cam_ns = leo_host::sec2nano(60.0);                                   // OND:62
    if (cam_ns % p.dyn_ns) return "the 60 s camera period must be a multiple of dynRate";
    p.cam_ticks = (int32_t)(cam_ns / p.dyn_ns);
    // hub (OND:176-185) and the HR16 pyramid (OND:269-293; [BSK: simIncludeRW.Honeywell_HR16], maxMomentum 50)
    p.I[0] = 900.; p.I[4] = 800.; p.I[8] = 600.;
    p.Om_max = 6000.0 * RPM; p.u_max = 0.2; p.Js = 50. / p.Om_max; p.invJs = 1.0 / p.Js;
    double D[9], M[9] = {0}, Mi[9];
    memcpy(D, p.I, sizeof(D));
    for (int i = 0; i < ON_NRW; i++) {
        const double el = 40.0 * D2R, az = (45.0 + 90.0 * i) * D2R;
        p.gs[i][0] = cos(el) * cos(az); p.gs[i][1] = cos(el) * sin(az); p.gs[i][2] = sin(el);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) { D[3 * a + b] -= p.Js * p.gs[i][a] * p.gs[i][b]; M[3 * a + b] += p.gs[i][a] * p.gs[i][b]; }
    }
    if (!leo_host::inv3(D, p.Dinv) || !leo_host::inv3(M, Mi)) return "singular wheel geometry";
    for (int i = 0; i < ON_NRW; i++)
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < 3; a++) p.Umap[i][b] += p.gs[i][a] * Mi[3 * a + b];
    p.mu_dyn = 4.2828371901284001E+13;                 // OND:386
    p.mu_fsw = 42828.314 * 1E9;                        // [BSK: astroConstants MU_MARS] * 1e9 (also ONF:507)
    p.K = 3.5; p.Pgain = 30.0;                          // ONF:400-402 (Ki = -1: integral off)
    // sigma_R0R = C2MRP(euler1(90) euler2(90) MRP2C(0)) = (1/3, 1/3, -1/3) (ONF:350-355)
    p.sigma_RR0[0] = -(0.5 / 1.5); p.sigma_RR0[1] = -(0.5 / 1.5); p.sigma_RR0[2] = 0.5 / 1.5;
    { const double n[ON_NCSS][3] = {{0.0, 0.707107, 0.707107}, {0.707107, 0., 0.707107}, {0.0, -0.707107, 0.707107},
                                    {-0.707107, 0., 0.707107}, {0.0, -0.965926, -0.258819}, {-0.707107, -0.353553, -0.612372},
                                    {0., 0.258819, -0.965926}, {0.707107, -0.353553, -0.612372}};   // OND:341-350
      memcpy(p.cssN, n, sizeof(n)); }
    p.css_cos_fov = cos(80. * D2R); p.css_scale = 2.0;  // OND:337-338
    p.R_sun = 695000.0 * 1000; p.R_planet = 3396.19 * 1000;   // [BSK: astroConstants REQ_SUN, REQ_MARS]
    p.inv_RsPlusRp = 1.0 / (p.R_sun + p.R_planet); p.inv_RsMinusRp = 1.0 / (p.R_sun - p.R_planet);
    { // simple_nav PMatrix diagonal and walk bounds (OND:238-253); the DV states are not used
        const double P[15] = {10.0, 10.0, 10.0, 0.001, 0.001, 0.001,
                              1.0 / 36000.0 * PI / 180.0, 1.0 / 36000.0 * PI / 180.0, 1.0 / 36000.0 * PI / 180.0,
                              0.00005 * PI / 180.0, 0.00005 * PI / 180.0, 0.00005 * PI / 180.0,
                              0.1 * PI / 180.0, 0.1 * PI / 180.0, 0.1 * PI / 180.0};
        const double B[15] = {100000.0, 100000.0, 100000.0, 0.1, 0.1, 0.1,
                              1E-18 * PI / 180.0, 1E-18 * PI / 180.0, 1E-18 * PI / 180.0,
                              1E-18 * PI / 180.0, 1E-18 * PI / 180.0, 1E-18 * PI / 180.0,
                              5.0 * PI / 180.0, 5.0 * PI / 180.0, 5.0 * PI / 180.0};
        memcpy(p.navP, P, sizeof(P)); memcpy(p.navBound, B, sizeof(B));
    }
    p.nav_noise = c.nav_noise ? 1 : 0; p.camera_reenable = c.camera_reenable ? 1 : 0;
    { // camera 512 x 512, FOV 55 deg (OND:138-141); pixelLine's normalised pixel pitch
        const double res = 512.0, fov = 55.0 * PI / 180.0, pX = 2. * tan(fov * res / res / 2.0);
        p.cam_res = res; p.cam_half = pX / 2; p.cam_X = pX / res;
    }
    p.hough_min_radius = 20.0;                          // ONF:463
    p.pixel_noise_std = c.pixel_noise_std; p.circle_unc = c.circle_unc;
    p.planet_radius_km = 3396.19;
    { // relativeODuKF: alpha 0.02, beta 2, kappa 0 (ONF:499-501); qNoise, noiseSF as overridden at ONS:196-200
        const double alpha = 0.02, n = 6.0, lambda = alpha * alpha * n - n;
        p.ukf_gamma = sqrt(n + lambda);
        p.ukf_w = 1.0 / 2.0 * 1.0 / (n + lambda); p.ukf_sqrt_w = sqrt(p.ukf_w);
        p.ukf_cm = 2.0 - alpha * alpha; p.ukf_sqrt_cm = sqrt(p.ukf_cm);
        p.ukf_sq_pos = sqrt(1E-3 * 1E-3); p.ukf_sq_vel = sqrt(1E-4 * 1E-4);
        p.ukf_noiseSF = 5.0;
        p.ukf_P0_pos = sqrt(1. * 1E6); p.ukf_P0_vel = sqrt(0.02 * 1E6);     // ONF:514-519
    }
    p.reward_mult = c.reward_mult;
    p.max_length = c.max_length; p.numModes = c.numModes; p.auto_reset = c.auto_reset; p.sample_orbit = c.sample_orbit;
    elem2rv(p.mu_dyn, 18000 * 1E3, 0.6, 10 * D2R, 25. * D2R, 190. * D2R, 80. * D2R, p.rN0, p.vN0);   // ONS:173-181
    // '2019 DECEMBER 12 18:00:00.0' UTC (OND:396): JD 2458830.25 -> 7285.25 days from J2000, + 69.184 s to TT
    p.epoch_days = 7285.25 + 69.184 / 86400.0;
    p.seed = c.noise_seed;
    return "";
}

// FP64 flop per env-decision-step of the opNav kernel AS BUILT (FMA = 2, add/mul = 1, one per MUFU seed), from the
// operation list of opnav_core.cuh and matched to ncu's executed 2*DFMA + DMUL + DADD thread-instruction counters
// (profiles/ncu_opnav_r01f.md: 5.774e11 per 32768-env launch = 5874 per tick and env with the 50/50 action mix):
//   filter time update   13 two-body RK4 steps (4 x 44 + 6), sigma points and deviations 6 x 60, Gram matrix 6 x 84,
//                        covariance assembly 100, 6 x 6 Cholesky 185                                        ~ 3400
//   truth RK4            4 x eom (gravity 20, wheel momentum and torque 52, gyroscopics + inverse 45, MRP kinematics 40,
//                        wheel rates 28) + stage combinations 4 x 64                                        ~  950
//   simple_nav           16 Box-Muller normals (atanh-series log, rsqrt, octant sincos: ~85 per pair) and 15 bounded-walk
//                        states (reciprocal, exp)                                                           ~ 1060
//   guidance + control   hillPoint / tracking error / MRP feedback / torque map (OpNav pointing) or eclipse / CSS /
//                        cssWlsEst / sunSafePoint / MRP feedback (sun-safe)                                  ~  400
//   nav message          MRP composition of the attitude error, position / velocity / rate sums             ~   55
// The measurement update (once per 60 ticks while imaging) adds ~10 per tick on average.
static inline double flops_per_step(const OpNavParams &p)
{
    const double per_tick = 3400.0 + 950.0 + (p.nav_noise ? 1060.0 : 0.0) + 400.0 + 55.0;
    return per_tick * p.ticks_per_step;
}

}  // namespace opnav_host
