// hostcore_opnav.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the opNav device core (basilisk_env_b200/csrc/opnav_core.cuh) for the HOST with g++ so that the fused
// schedule and the streaming square-root filter can be compared with the independent oracle on a CPU-only box
// (pytest -m "not gpu").  Never loaded by the product: libbskenv.so has no CPU path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../basilisk_env_b200/csrc/opnav_core.cuh"
#include "../../basilisk_env_b200/csrc/opnav_host.h"

struct HCO {
    OpNavParams P;
    int64_t n;
    std::vector<double> S, ics;
    std::vector<int64_t> I;
    std::vector<double> eph;
};

extern "C" {
void hco_set_ephemeris(struct HCO *h, double t0, double seg_len, int nseg, int ncoef, const double *coef);
void hco_default_config(bskenv_opnav_config *c) { opnav_host::default_config(c); }
HCO *hco_create(const bskenv_opnav_config *cfg, int64_t n, int64_t first_env)
{
    HCO *h = new HCO();
    std::string err = opnav_host::build_params(*cfg, h->P);
    if (!err.empty()) { delete h; return nullptr; }
    h->P.first_env_index = first_env;
    h->n = n;
    h->S.assign((size_t)OPNAV_ND * n, 0.0);
    h->I.assign((size_t)OPNAV_NI * n, 0);
    h->ics.assign((size_t)OPNAV_IC_DIM * n, 0.0);
    return h;
}
void hco_destroy(HCO *h) { delete h; }
void hco_reset_ics(HCO *h, const double *ics, double *obs)
{
    for (int64_t e = 0; e < h->n; e++) {
        double ic[OPNAV_IC_DIM];
        for (int k = 0; k < OPNAV_IC_DIM; k++) { ic[k] = ics[e * OPNAV_IC_DIM + k]; h->ics[e * OPNAV_IC_DIM + k] = ic[k]; }
        opnav::opnav_reset_env(h->P, h->S.data(), h->I.data(), h->n, e, ic, obs ? obs + 4 * e : nullptr);
    }
}
void hco_reset_seeded(HCO *h, uint64_t seed, double *ics_out, double *obs)
{
    h->P.seed = seed;
    for (int64_t e = 0; e < h->n; e++) {
        double ic[OPNAV_IC_DIM];
        int64_t ep = h->I[(size_t)OI_EPISODE * h->n + e] + 1;
        h->I[(size_t)OI_EPISODE * h->n + e] = ep;
        opnav::sample_ic(h->P, h->P.first_env_index + e, ep, ic);
        for (int k = 0; k < OPNAV_IC_DIM; k++) { h->ics[e * OPNAV_IC_DIM + k] = ic[k]; if (ics_out) ics_out[e * OPNAV_IC_DIM + k] = ic[k]; }
        opnav::opnav_reset_env(h->P, h->S.data(), h->I.data(), h->n, e, ic, obs ? obs + 4 * e : nullptr);
    }
}
void hco_step(HCO *h, const int32_t *actions, double *obs, double *reward, uint8_t *done, uint8_t *reason, double *debug)
{
    for (int64_t e = 0; e < h->n; e++) {
        opnav::StepOut o;
        opnav::Ukf f;
        opnav::Cold c;
        opnav::Walk w;
        std::vector<double> mbuf((size_t)opnav::opnav_meas_slots(h->P) * ON_MEAS_W);
        opnav::MeasBuf mb;
        mb.p = mbuf.data(); mb.stride = 1;
        if (getenv("HCO_NOISE_SPLIT") && getenv("HCO_NOISE_SPLIT")[0] == '1') {
            // the three-kernel form of the interval (opnav.cu: opnav_noise_kernel / opnav_dyn_kernel / opnav_pass2_kernel), back to back
            std::vector<double> nz((size_t)(h->P.ticks_per_step + 1) * 15);
            double feed[30];
            opnav::opnav_pass0(h->P, h->S.data(), h->I.data(), h->n, e, w, nz.data(), 1);
            opnav::NoiseFeed nf;
            nf.g = nz.data(); nf.stride = 1; nf.buf = feed;
            opnav::opnav_pass1_fed(h->P, h->S.data(), h->I.data(), h->n, e, actions[e], c, nf, mb);
            opnav::opnav_pass2(h->P, h->S.data(), h->I.data(), h->n, e, actions[e], o, f, mb);
        } else
        opnav::opnav_step_env(h->P, h->S.data(), h->I.data(), h->n, e, actions[e], o, f, c, w, mb);
        for (int k = 0; k < 4; k++) obs[4 * e + k] = o.ob[k];
        if (debug) for (int k = 0; k < 12; k++) debug[12 * e + k] = o.debug[k];
        reward[e] = o.reward; done[e] = (uint8_t)o.done; reason[e] = (uint8_t)o.reason;
    }
}
void hco_get_state(HCO *h, double *S, int64_t *I)
{
    memcpy(S, h->S.data(), h->S.size() * sizeof(double));
    memcpy(I, h->I.data(), h->I.size() * sizeof(int64_t));
}
int hco_dims(int *nd, int *ni) { *nd = OPNAV_ND; *ni = OPNAV_NI; return 0; }
void hco_normals(HCO *h, int64_t env, int64_t episode, uint32_t tick, uint32_t stream, uint32_t block, double *out)
{
    double n4[4];
    opnav::normals4(h->P, env, episode, tick, stream, block, n4);
    for (int k = 0; k < 4; k++) out[k] = n4[k];
}
void hco_sun(HCO *h, double t, double *r, double *v)
{
    opnav::SunState s = opnav::sun_from_mars(h->P, t);
    r[0] = s.r.x; r[1] = s.r.y; r[2] = s.r.z; v[0] = s.v.x; v[1] = s.v.y; v[2] = s.v.z;
}
double hco_eclipse(HCO *h, const double *sun, const double *r)
{
    return opnav::eclipse_mars(h->P, leo::mk(sun[0], sun[1], sun[2]), leo::mk(r[0], r[1], r[2]));
}
// streaming SR-UKF on caller data: x[6], S[21] lower triangle row-major
void hco_ukf_time_update(HCO *h, double *x, double *S, double *m, double dt)
{
    opnav::Ukf f;
    for (int i = 0; i < 6; i++) { f.x[i] = x[i]; f.m[i] = 0; }
    for (int r = 0, k = 0; r < 6; r++) for (int c = 0; c <= r; c++) f.SC(r, c) = S[k++];
    (void)opnav::ukf_time_update(h->P, f, dt);
    for (int i = 0; i < 6; i++) { x[i] = f.x[i]; m[i] = f.m[i]; }
    for (int r = 0, k = 0; r < 6; r++) for (int c = 0; c <= r; c++) S[k++] = f.SC(r, c);
}
int hco_ukf_meas_update(HCO *h, double *x, double *S, const double *m, double dt, const double *obs, const double *R6)
{
    opnav::Ukf f;
    for (int i = 0; i < 6; i++) { f.x[i] = x[i]; f.m[i] = m[i]; }
    for (int r = 0, k = 0; r < 6; r++) for (int c = 0; c <= r; c++) f.SC(r, c) = S[k++];
    double o[3] = {obs[0], obs[1], obs[2]}, R[6];
    for (int i = 0; i < 6; i++) R[i] = R6[i];
    bool ok = opnav::ukf_meas_update(h->P, f, dt, o, R);
    for (int i = 0; i < 6; i++) x[i] = f.x[i];
    for (int r = 0, k = 0; r < 6; r++) for (int c = 0; c <= r; c++) S[k++] = f.SC(r, c);
    return ok ? 1 : 0;
}
void hco_set_ephemeris(HCO *h, double t0, double seg_len, int nseg, int ncoef, const double *coef)
{   // the setting bskenv_opnav_set_ephemeris offers
    h->eph.assign(coef, coef + (size_t)(nseg > 0 ? nseg : 0) * 3 * ncoef);
    LeoEph &E = h->P.eph_sun;
    E.coef = nseg > 0 ? h->eph.data() : nullptr; E.nseg = nseg > 0 ? nseg : 0; E.ncoef = ncoef; E.t0 = t0; E.seg_len = seg_len;
}
}
