/*
 * opnav_oracle.h -- public interface of the CPU oracle for the opNav environment
 * (dynamics half + synthetic nav measurement + relative-OD filter; no Vizard rendering).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic lives in AVS-Lab Basilisk 1.x (unpinned, inferred 1.8.x), absent from
 * /root/reference and unbuildable here; the reference ships no golden outputs.  This is a scalar FP64
 * restatement of the Basilisk 1.x module algorithms wired as
 *   /root/reference/basilisk_env/simulators/opNavSimulator.py                  (ONS:line)
 *   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavDynamics.py  (OND:line)
 *   /root/reference/basilisk_env/simulators/opNav_models/BSK_OpNavFsw.py       (ONF:line)
 *   /root/reference/basilisk_env/simulators/opNav_models/BSK_masters.py        (ONM:line)
 *   /root/reference/basilisk_env/envs/opNavEnvironment.py                      (ONE:line)
 * wire them.  The camera -> Vizard -> houghCircles image path is replaced by a pinhole projection of the
 * Mars disc (BASELINE north_star: "nav measurement model batched (no Vizard rendering)").
 */
#ifndef OPNAV_ORACLE_H
#define OPNAV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-env initial conditions (ONS:163-202): truth orbit and the filter's initial state error. */
typedef struct {
    double rN[3], vN[3];       /* elem2rv(mu_mars, a=18000 km, e=0.6, i=10, Om=25, om=190, f=80 deg) (ONS:173-181) */
    double rError[3];          /* uniform(100000,-100000,3) m   (ONS:187) */
    double vError[3];          /* uniform(1000,-1000,3) m/s     (ONS:188) */
} orc_opnav_ic;

typedef struct {
    double dynRate;            /* 1.0 s  (ONE:86: scenario_OpNav(1., 1., step_duration)) */
    double fswRate;            /* 1.0 s */
    double step_duration_min;  /* 50.  minutes: simTime += 50 then mc.min2nano (ONS:256-257) */
    int    nav_noise;          /* 1: simple_nav Gauss-Markov errors with the OND:236-258 PMatrix/walkBounds; 0: truth */
    int    camera_reenable;    /* 0: reference behaviour -- action 1 clears cameraIsOn and nothing sets it again
                                  (ONS:239 is commented out); 1: action 0 switches the camera back on */
    double pixel_noise_std;    /* synthetic circle finder: 1-sigma noise on centre x, y and radius [px] */
    double circle_unc;         /* synthetic circle finder: CirclesOpNavMsg.uncertainty diagonal [px^2] */
    uint64_t seed;             /* key of the per-env noise streams (counter-based, see opnav_oracle.c) */
    int    numModes;           /* 50 (ONS:149) */
    int    reserved[3];
} orc_opnav_cfg;

typedef struct {
    double r_BN_N[3], v_BN_N[3], sigma_BN[3], omega_BN_B[3], Omega[4], u_current[4];
    double navErrors[18];
    double nav_r[3], nav_v[3], nav_sigma[3], nav_omega[3], nav_sun_B[3];
    double sigma_BR[3], omega_BR_B[3], Lr[3], rwCmd[4];
    double css[8], sun_point[3], shadow;
    double filt_state[6], filt_covar[36], filt_sBar[36], filt_time;
    double meas_r[3], meas_covar[9];
    double circle[3];
    int64_t n_meas, n_bad, n_images, mrp_switch_count;
    int32_t camera_on, mode, modeCounter, pad;
    uint64_t sim_nanos;
} orc_opnav_state;

typedef struct orc_opnav_sim orc_opnav_sim;

void orc_opnav_default_cfg(orc_opnav_cfg *cfg);
/* reference ICs (fixed orbit); rError/vError left zero */
void orc_opnav_reference_orbit(orc_opnav_ic *ic);
orc_opnav_sim *orc_opnav_create(const orc_opnav_ic *ic, const orc_opnav_cfg *cfg, uint64_t env_index, uint64_t episode);
void orc_opnav_destroy(orc_opnav_sim *s);
/* run_sim(action) (ONS:225-299): obs[4], debug[12]; returns sim_over */
int orc_opnav_run_sim(orc_opnav_sim *s, int action, double obs[4], double debug[12]);
void orc_opnav_get_state(const orc_opnav_sim *s, orc_opnav_state *out);

/* gym layer (ONE:55-125) */
typedef struct { double ob[4]; double debug[12]; double reward; int done; int reason; } orc_opnav_out;
typedef struct orc_opnav_env orc_opnav_env;
orc_opnav_env *orc_opnav_env_create(const orc_opnav_cfg *cfg);
void orc_opnav_env_destroy(orc_opnav_env *e);
void orc_opnav_env_reset(orc_opnav_env *e, const orc_opnav_ic *ic, uint64_t env_index, uint64_t episode, double ob[4]);
void orc_opnav_env_step(orc_opnav_env *e, int action, orc_opnav_out *out);
orc_opnav_sim *orc_opnav_env_sim(orc_opnav_env *e);
void orc_opnav_env_set_max_length(orc_opnav_env *e, int max_length);
void orc_opnav_env_episode(const orc_opnav_env *e, double *reward_total, int *curr_step_at_info);
void orc_opnav_env_step_batch(orc_opnav_env **envs, int n, const int *actions, orc_opnav_out *outs, int nthreads);

/* helpers exposed for unit tests */
void orc_opnav_normals(uint64_t seed, uint64_t env, uint64_t episode, uint32_t tick, uint32_t stream, uint32_t block,
                       double out[4]);
/* pinhole projection of a sphere of radius R at r_C (camera frame, +z boresight): centre px (x, y), radius px; returns valid */
int orc_opnav_project_circle(const double r_planet_C[3], double planet_radius, double circle[3]);
/* pixelLineConverter: circle + uncertainty diag + dcm_NC -> r_BN_N [m], covar_N [m^2] */
void orc_opnav_pixel_line(const double circle[3], double unc, double dcm_CN[3][3], double r_BN_N[3], double covar_N[9]);
/* SR-UKF pieces on caller-provided arrays (n = 6 states, 3 obs) */
void orc_ukf_qr_just_r(const double *A, int nRow, int nCol, double *R);
int  orc_ukf_chol_downdate(const double *rMat, const double *xVec, double beta, int n, double *rOut);
int  orc_ukf_chol_decomp(const double *A, int n, double *L);
void orc_ukf_state_prop(double x[6], double mu, double dt);
typedef struct {
    double state[6], sBar[36], covar[36], xBar[6], SP[13 * 6], timeTag;
    double wM[13], wC[13], gamma, sQnoise[36], mu, noiseSF;
    int64_t n_bad;
} orc_ukf;
void orc_ukf_init(orc_ukf *f, const double state[6], const double covar[36], const double qNoise[36], double mu, double noiseSF);
void orc_ukf_time_update(orc_ukf *f, double updateTime);
void orc_ukf_meas_update(orc_ukf *f, const double obs[3], const double covar_N[9]);

#ifdef __cplusplus
}
#endif
#endif
