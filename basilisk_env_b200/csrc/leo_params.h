// leo_params.h -- kernel parameter block and persistent-state layout of the fused LEO step.
//
// LeoParams is passed BY VALUE as a __grid_constant__ kernel parameter: every field is then a
// constant-bank operand of the FP64 instructions (no registers, no loads).  It is derived on the
// host from bskenv_config (include/bskenv.h) by leo_build_params(); the numbers come from
//   /root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py:127-191, 195-490 (SIM)
//   /root/reference/basilisk_env/simulators/dynamics/effectorPrimatives/actuatorPrimatives.py (AP)
//   /root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:25-42 (ENV)
#pragma once
#include <stdint.h>

#define LEO_MAX_RW 4
#define LEO_NTHR 8
#ifndef LEO_BLOCK
#define LEO_BLOCK 128          // threads per block of the step kernel = stride of the shared-memory message bus
#endif

// Chebyshev ephemeris table in the layout of SPICE SPK type 2 / binary PCK type 2 records: nseg segments
// of equal length from t0 [s of sim time], three components x ncoef coefficients each, coef[nseg][3][ncoef] in device
// memory (every thread of a warp reads the same address: one broadcast transaction).  nseg = 0: no table.
struct LeoEph { const double *coef; double t0, seg_len; int32_t nseg, ncoef; };

struct LeoParams {
    // ---- task rates as integer nanoseconds (Basilisk sec2nano) and derived loop counts ----
    int64_t dyn_ns, fsw_ns, step_ns;
    int32_t ticks_per_fsw, fsw_per_step;
    // ---- hub ----
    double inv_mass;
    double I[9];          // IHubPntBc_B (= ISCPntB_B: balanced wheels carry no mass properties)
    double D[9];          // I - sum_i Js_i g_i g_i^T   (constant back-substitution matrix of the balanced wheels)
    double Dinv[9];       // its inverse
    // ---- gravity ----
    double mu_c, mu_sun;
    double j2k;           // 1.5 * J2 * mu * Req^2 (only when the J2 template flag is on)
    int32_t use_j2, hill_cel_pun;
    // degree-2 field in the planet-fixed frame (bskenv_set_gravity_degree2).  gm2 = mu Req^2 [M] with
    // U_2 = mu Req^2 (r^T M r) / |r|^5 for the un-normalised C20, C21, S21, C22, S22: xx, yy, zz, xy, xz, yz
    int32_t grav_pfix, pad3;
    double gm2[6];
    LeoEph eph_sun;       // Sun position relative to Earth [m] (replaces the analytic Sun, deviation D1)
    LeoEph eph_orient;    // Earth orientation angles RA, DEC, W [rad] (replaces the IAU rotation model)
    int32_t mixed, pad2;   // 1: FP32 stage arithmetic with FP64 accumulation (leo_f32.cuh); stress-config trade-off only
    int32_t diag, pad1;    // 1: diagonal hub inertia, three wheels along the body axes, drag facets on their normal axis
                           //    (the reference set-up): fast EOM path
    // ---- reaction wheels ----
    int32_t nrw, pad0;
    double gs[LEO_MAX_RW][3], Js[LEO_MAX_RW], invJs[LEO_MAX_RW];
    double u_max[LEO_MAX_RW], u_min[LEO_MAX_RW], Om_max[LEO_MAX_RW];
    double Umap[LEO_MAX_RW][3];   // rwMotorTorque: us = Umap * (-Lr) = CGs^T (CGs CGs^T)^-1 C (-Lr)
    // ---- facet drag, facets with axis-aligned normals collapsed per axis and sign ----
    // K(+/-) = sum of 0.5*Cd*A over the facets with normal (+/-) e_axis, M(+/-) = sum of 0.5*Cd*A*r_facet over the
    // same facets; a facet acts when v_B has a positive component along its normal, so the selected sums times
    // |v_axis| are  Ka |v| + Kd v  with  Ka = (K+ + K-)/2,  Kd = (K+ - K-)/2  (no table look-up, no branch).
    // dragKa / dragKd are pre-divided by the spacecraft mass.
    double dragKa[3], dragKd[3];
    double dragMa[3][3], dragMd[3][3];   // [axis][component]
    double dist_mag;              // disturbance_magnitude (2e-4)
    // ---- exponential atmosphere ----
    double rho0, inv_H, Rp_atmo;
    // ---- eclipse ----
    double R_sun, R_planet;
    // ---- power ----
    double nHat_B[3];
    double panel_coef;            // eta * SOLAR_FLUX_EARTH * A * AU^2
    double sink_power, capacity;
    // ---- FSW ----
    double K, P, Ki;
    double I_fsw[9];
    double sigma_R0N[3];
    double hs_min;
    double thr_loc[LEO_NTHR][3], thr_dir[LEO_NTHR][3];
    double thr_D[3][LEO_NTHR];    // r_i x t_i
    double thr_W[LEO_NTHR][3];    // D^T (D D^T)^-1  (minimum-norm map, thrForceMapping)
    double thr_Fmax, thr_MinOnTime, thrMinFireTime;
    double tfm_eps, tfm_angErrThresh;
    int32_t thrForceSign, maxCounterValue;
    // ---- gym layer ----
    double wheel_limit, power_max, reward_mult, failure_penalty, decay_radius, wheel_rpm2rad;
    int32_t max_length, auto_reset;
    // ---- epoch (days of TT from J2000 at sim time 0) ----
    double epoch_days;
    // ---- IC sampling ----
    uint64_t seed;
    int64_t first_env_index;
};

// FP32 copies of the parameters the mixed-precision tick reads (leo_f32.cuh); second __grid_constant__ kernel argument
struct LeoParamsF {
    float mu_c, mu_sun, j2k;
    float D[9], Dinv[9];
    float dragKa[3], dragKd[3], dragMa[3][3], dragMd[3][3];
    float rho0, inv_H, Rp_atmo, R_sun, R_planet;
    float nHat_B[3], panel_coef;
};

// ---- persistent per-env state: double fields (SoA, field-major, stride = padded env count) ----
enum LeoDField : int {
    F_R = 0,         // r_BN_N[3]
    F_V = 3,         // v_BN_N[3]
    F_SIG = 6,       // sigma_BN[3]
    F_OMG = 9,       // omega_BN_B[3]
    F_WHL = 12,      // Omega[4]
    F_UCUR = 16,     // RW u_current[4]
    F_RHO = 20,      // latched neutral density
    F_E = 21,        // storedCharge
    F_SHADOW = 22,   // shadowFactor of the last env tick
    F_LDIST = 23,    // extTorquePntB_B[3]
    F_GUID = 26,     // att_guidance: sigma_BR, omega_BR_B, omega_RN_B, domega_RN_B
    F_REF = 38,      // att_reference: sigma_RN, omega_RN_N, domega_RN_N
    F_LR = 47,       // commandedControlTorque[3]
    F_RWCMD = 50,    // rwTorqueCommand[4]
    F_DELTAH = 54,   // wheelDeltaH[3]
    F_THRON = 57,    // ThrustOnCmd[8]
    F_THRSTART = 65, // ThrusterStartTime (shared by all thrusters of a command)
    F_THRPREVFIRE = 66,
    F_THRREM = 67,   // thrOnTimeRemaining[8]
    F_THRCMD = 75,   // rwDesatTimeOnCmd OnTimeRequest[8]
    F_EPRET = 83,    // reward_total of the running episode
    F_OBS = 84,      // last un-normalised simulator obs[5]
    LEO_ND = 89
};
// ---- persistent per-env state: int64 fields ----
enum LeoIField : int {
    I_TICK = 0,      // index of the last executed dynamics tick (-1 after reset)
    I_STEP = 1,      // curr_step
    I_MASK = 2,      // active tasks: 1 sunPoint, 2 nadirPoint, 4 mrpControl, 8 rwDesat
    I_SWITCH = 3,    // MRPSwitchCount
    I_THRFACTOR = 4, // bit i: ThrustFactor_i > 0
    I_INITREQ = 5,   // thrMomentumManagement.initRequest
    I_DUMPCNT = 6,   // thrMomentumDumping.thrDumpingCounter
    I_DUMPPRIOR = 7, // thrMomentumDumping.priorTime [ns]
    I_LASTDH = 8,    // thrMomentumDumping.lastDeltaHInMsgTime [ns]
    I_DHTIME = 9,    // write time of wheelDeltaH [ns]; 0 when never written
    I_EPISODE = 10,  // episode counter (keys the IC stream)
    I_FIRE = 11,     // fireCounter[8]
    I_THRACTIVE = 19,// any ThrustOnCmd > 0 or ThrustFactor > 0
    I_OVER = 20,     // episode_over latch
    I_RWSAT = 21,    // a wheel was speed-saturated at the last latch
    LEO_NI = 22
};

#define LEO_TASK_SUN 1
#define LEO_TASK_NADIR 2
#define LEO_TASK_MRP 4
#define LEO_TASK_DESAT 8
#define LEO_TASK_ALL 15
