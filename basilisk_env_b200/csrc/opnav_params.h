// opnav_params.h -- kernel parameter block and persistent-state layout of the fused opNav step
// (dynamics half of the opNav env + synthetic nav measurement + relative-OD SR-UKF; no rendering).
//
// Numbers come from
//   /root/reference/basilisk_env/simulators/opNavSimulator.py                  (ONS:line)
//   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavDynamics.py  (OND:line)
//   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavFsw.py       (ONF:line)
//   /root/reference/basilisk_env/envs/opNavEnvironment.py                      (ONE:line)
#pragma once
#include <stdint.h>
#include "leo_params.h"      // LeoEph: Chebyshev ephemeris table

#define ON_NRW 4
#define ON_NCSS 8
#ifndef ON_BLOCK
#define ON_BLOCK 128
#endif

struct OpNavParams {
    int64_t dyn_ns;                 // dynRate = fswRate as integer nanoseconds
    int32_t ticks_per_step;         // step_duration [min] * 60 / dynRate  (3000)
    int32_t cam_ticks;              // CameraTask period in ticks (60 s, OND:62)
    double dt;                      // dynRate [s]
    // ---- hub + wheels (OND:176-185, :269-293) ----
    double I[9], Dinv[9];           // hub inertia; inverse of I - sum Js g g^T
    double gs[ON_NRW][3], Js, invJs, u_max, Om_max;
    double Umap[ON_NRW][3];         // rwMotorTorque: us = Umap (-Lr)
    double mu_dyn, mu_fsw;          // OND:386 vs [BSK astroConstants MU_MARS] used by the filter
    // ---- FSW (ONF:345-356, :399-409) ----
    double K, Pgain;
    double sigma_RR0[3];            // -sigma_R0R of trackingErrorCam
    double cssN[ON_NCSS][3], css_cos_fov, css_scale;
    double R_sun, R_planet, inv_RsPlusRp, inv_RsMinusRp;
    // ---- simple_nav (OND:236-258) ----
    double navP[15], navBound[15];
    int32_t nav_noise, camera_reenable;
    // ---- synthetic camera / circle finder / pixelLine (OND:129-143, ONF:452-475) ----
    double cam_X, cam_half, cam_res, hough_min_radius, pixel_noise_std, circle_unc, planet_radius_km;
    // ---- relativeODuKF (ONF:495-527, ONS:196-200) ----
    double ukf_gamma, ukf_w, ukf_sqrt_w, ukf_cm, ukf_sqrt_cm;   // w = 1/(2(n+lambda)); cm = 2 - alpha^2 (see opnav_core.cuh)
    double ukf_sq_pos, ukf_sq_vel, ukf_noiseSF;
    double ukf_P0_pos, ukf_P0_vel;                              // sqrt of the initial covariance diagonal
    // ---- gym layer (ONE:23-44, ONS:149) ----
    double reward_mult;
    int32_t max_length, numModes, auto_reset, sample_orbit;
    // ---- reference orbit (ONS:173-181) ----
    double rN0[3], vN0[3];
    // ---- epoch: TT days from J2000 at sim time 0 ('2019 DECEMBER 12 18:00:00.0', OND:396) ----
    double epoch_days;
    LeoEph eph_sun;                 // optional table: Sun position relative to the Mars barycentre [m] (bskenv_opnav_set_ephemeris)
    uint64_t seed;                  // noise / IC stream key
    int64_t first_env_index;
};

// ---- persistent per-env state: double fields (SoA, field-major, stride = padded env count) ----
enum OpNavDField : int {
    OF_R = 0, OF_V = 3, OF_SIG = 6, OF_OMG = 9, OF_WHL = 12,   // truth states r, v, sigma, omega, Omega[4]
    OF_RWCMD = 16,      // reactionwheel_cmds motorTorque[4] (written by FSW, latched by the effector one tick later)
    OF_NAVERR = 20,     // simple_nav Gauss-Markov error states [15]
    OF_SUNPT = 35,      // "sun_point_data" vehSunPntBdy[3] (cssWlsEst output of the previous sun-safe pass)
    OF_SHADOW = 38,     // eclipse message of the previous tick
    OF_FSTATE = 39,     // filter state [6]
    OF_FS = 45,         // filter sBar, lower triangle row-major [21]
    OF_EPRET = 66,      // reward_total
    OF_OBS = 67,        // last obs[4]
    OF_DEBUG = 71,      // last debug states [12]: nav r, true r, true v, sigma_BN
    OPNAV_ND = 83
};
enum OpNavIField : int {
    OI_TICK = 0,        // index of the last executed tick (-1 after reset)
    OI_STEP = 1,        // curr_step
    OI_MODE = 2,        // FSW task set: 0 OpNavOD (opNavPointTaskCheat + mrpFeedbackRWs + opNavOD), 1 sunSafePoint + mrpFeedbackRWs
    OI_CAMERA = 3,      // cameraMod.cameraIsOn
    OI_MODECNT = 4,     // modeCounter
    OI_FIRST = 5,       // 1 until the first run_sim (pending 'OpNavOD' event)
    OI_SWITCH = 6,      // MRPSwitchCount
    OI_EPISODE = 7,     // episode counter (keys the noise and IC streams)
    OI_OVER = 8,        // episode_over latch
    OI_NMEAS = 9,       // measurement updates
    OI_NBAD = 10,       // rejected filter updates
    OI_FTICK = 11,      // filter time tag in ticks
    OI_SUNPT_W = 12,    // sun_point_data written
    OI_NIMG = 13,       // frames taken
    OPNAV_NI = 14
};
#define OPNAV_IC_DIM 12
#define OPNAV_OBS_DIM 4
#define OPNAV_DEBUG_DIM 12
