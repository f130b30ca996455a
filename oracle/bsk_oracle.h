/*
 * bsk_oracle.h -- public interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in AVS-Lab Basilisk 1.x
 * (unpinned, inferred 1.8.x), which is not in /root/reference and cannot be built in this image.
 * The reference itself ships no tests, golden vectors or recorded trajectories.  This oracle is a
 * scalar FP64 restatement of the Basilisk 1.x module algorithms wired exactly as
 * /root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py wires them.
 */
#ifndef BSK_ORACLE_H
#define BSK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Initial conditions of one LEO power/attitude sim: the per-env part of the reference's
 * `initial_conditions` dict (simulators/leoPowerAttitudeSimulator.py:127-191). */
typedef struct {
    double rN[3];            /* m     "rN"  (:133) */
    double vN[3];            /* m/s   "vN"  (:134) */
    double sigma_init[3];    /*       "sigma_init" (:142) */
    double omega_init[3];    /* rad/s "omega_init" (:143) */
    double disturbance_vector[3]; /* raw N(0,1)^3 draw (:152); torque = 2e-4 * this (:295, quirk Q4) */
    double wheelSpeeds_rpm[3];    /* RPM (:155, converted with mc.RPM at :303-305) */
    double storedCharge_Init;     /* W*s (:167) */
} orc_leo_ic;

/* Batch-global switches (deviations from the literal reference wiring are OFF by default). */
typedef struct {
    double dynRate;          /* 0.1  (envs/leoPowerAttitudeEnvironment.py:185) */
    double fswRate;          /* 1.0 */
    double step_duration;    /* 180. */
    int    use_j2;           /* 0: reference has no J2; 1: stress config */
    int    hill_cel_pun;     /* 0: planet at origin (intended); 1: type-punned SPICE msg */
    int    rw_set;           /* 0: balancedHR16Triad (actuatorPrimatives.py:20-37); 1: the four-wheel HR16 pyramid of
                                opNav_models/BSK_OpNavDynamics.py:269-293 (stress config); fourth wheel starts at the
                                mean of the three sampled speeds */
    int    grav_pfix;        /* 1 = degree-2 field (C20, C21, S21, C22, S22) evaluated in the planet-fixed
                                frame, orientation from the SPICE message Euler-stepped like the body positions */
    int    reserved[4];
} orc_leo_cfg;

/* Everything a parity test wants to see at a decision boundary. */
typedef struct {
    double r_BN_N[3], v_BN_N[3], sigma_BN[3], omega_BN_B[3], Omega[4];
    double u_current[4];
    double storedCharge, shadowFactor, density;
    double sigma_BR[3], omega_BR_B[3], sigma_RN[3];
    double Lr[3];
    double thrOnCmd[8];
    double thrOnTimeRemaining[8];
    double deltaH[3];
    double sun_r[3], sun_v[3];
    int64_t mrp_switch_count;
    int64_t thr_fire_count[8];
    int32_t thr_factor_mask;
    int32_t dump_counter;
    int32_t init_request;
    int32_t task_mask;       /* bit0 sunPoint, bit1 nadirPoint, bit2 mrpControl, bit3 rwDesat */
    uint64_t sim_nanos;
} orc_leo_state;

typedef struct orc_leo_sim orc_leo_sim;

void orc_leo_default_cfg(orc_leo_cfg *cfg);
orc_leo_sim *orc_leo_create(const orc_leo_ic *ic, const orc_leo_cfg *cfg);
void orc_leo_destroy(orc_leo_sim *s);
/* run_sim(action): mode switch, advance step_duration, sample the message log.
 * obs[5] = [|sigma_BR|, |omega_BN_B|, |Omega|, storedCharge/3600, shadowFactor]  (un-normalised,
 * simulators/leoPowerAttitudeSimulator.py:636-637); returns sim_over (:641-642). */
int orc_leo_run_sim(orc_leo_sim *s, int action, double obs[5]);
void orc_leo_initial_obs(const orc_leo_sim *s, double obs[5]);
void orc_leo_get_state(const orc_leo_sim *s, orc_leo_state *out);

/* The gym layer (envs/leoPowerAttitudeEnvironment.py:65-145) restated on top of run_sim. */
typedef struct {
    double ob[5];        /* normalised observation */
    double reward;
    int    done;
    int    reason;       /* bit0 max_length, bit1 wheel, bit2 power, bit3 decay */
} orc_env_out;
typedef struct orc_leo_env orc_leo_env;
orc_leo_env *orc_env_create(const orc_leo_cfg *cfg);
void orc_env_destroy(orc_leo_env *e);
void orc_env_reset(orc_leo_env *e, const orc_leo_ic *ic, double ob[5]);
void orc_env_step(orc_leo_env *e, int action, orc_env_out *out);
orc_leo_sim *orc_env_sim(orc_leo_env *e);
void orc_env_set_max_length(orc_leo_env *e, int max_length);
void orc_env_episode(const orc_leo_env *e, double *reward_total, int *curr_step_at_info);   /* ENV:130-136 */

/* Batched driver for the CPU baseline: steps n independent envs, OpenMP over envs. */
void orc_env_step_batch(orc_leo_env **envs, int n, const int *actions, orc_env_out *outs, int nthreads);
int  orc_max_threads(void);

/* Helpers exposed for unit tests (each cites its Basilisk counterpart in the .c file). */
void orc_elem2rv(double mu, double a, double e, double i, double Omega, double omega, double f,
                 double r[3], double v[3]);
void orc_sun_ephemeris(double t_sim_sec, double r[3], double v[3], double *j2000_et);
/* SURVEY 8(f)-4 (process-global test settings): Chebyshev ephemeris tables (kind 0: Sun position rel. Earth [m],
 * kind 1: Earth orientation angles RA, DEC, W [rad]; coef[nseg][3][ncoef]; nseg = 0 unloads) and the normalised
 * degree-2 coefficients C20, C21, S21, C22, S22. */
int orc_set_ephemeris(int kind, double t0, double seg_len, int nseg, int ncoef, const double *coef);
int orc_eph_eval(int kind, double t, double val[3], double rate[3]);
void orc_set_gravity_coeffs(const double cbar[5]);
void orc_earth_orientation(double t_sim_sec, double P[3][3], double Pdot[3][3]);
void orc_grav_degree2_pfix(double mu, double radEquator, const double cbar[5], const double pos[3], double acc[3]);
double orc_eclipse_shadow(const double r_sun[3], const double r_planet[3], const double r_sc[3], double planet_radius);
void orc_MRP2C(const double q[3], double C[3][3]);
void orc_C2MRP(double C[3][3], double q[3]);
void orc_subMRP(const double q1[3], const double q2[3], double out[3]);
void orc_addMRP(const double q1[3], const double q2[3], double out[3]);
void orc_thr_force_mapping(const double Lr[3], double F[8], double *angErr);

#ifdef __cplusplus
}
#endif
#endif
