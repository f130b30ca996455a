// hostcore.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the device core (basilisk_env_b200/csrc/leo_core.cuh) for the HOST with g++ so that the
// fused-schedule logic can be compared with the independent oracle on a CPU-only box (pytest -m "not gpu").
// It is never loaded by the product: the library proper (libbskenv.so) has no CPU path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../basilisk_env_b200/csrc/leo_core.cuh"
#include "../../basilisk_env_b200/csrc/leo_host.h"

struct HC {
    LeoParams P;
    LeoParamsF PF;
    int64_t n;
    std::vector<double> S, ics;
    std::vector<int64_t> I;
    int force_general = 0;
    std::vector<double> eph[2];
};

extern "C" {
void hc_default_config(bskenv_config *c) { leo_host::default_config(c); }
HC *hc_create(const bskenv_config *cfg, int64_t n)
{
    HC *h = new HC();
    std::string err = leo_host::build_params(*cfg, h->P);
    if (!err.empty()) { delete h; return nullptr; }
    leo_host::build_params_f(h->P, h->PF);
    h->n = n;
    h->S.assign((size_t)LEO_ND * n, 0.0);
    h->I.assign((size_t)LEO_NI * n, 0);
    h->ics.assign((size_t)19 * n, 0.0);
    return h;
}
void hc_destroy(HC *h) { delete h; }
void hc_reset_ics(HC *h, const double *ics, double *obs)
{
    for (int64_t e = 0; e < h->n; e++) {
        double ic[19];
        for (int k = 0; k < 19; k++) { ic[k] = ics[e * 19 + k]; h->ics[e * 19 + k] = ic[k]; }
        leo::leo_reset_env(h->P, h->S.data(), h->I.data(), h->n, e, ic, obs ? obs + 5 * e : nullptr);
    }
}
void hc_reset_seeded(HC *h, uint64_t seed, int64_t first_env, double *ics_out, double *obs)
{
    h->P.seed = seed; h->P.first_env_index = first_env;
    for (int64_t e = 0; e < h->n; e++) {
        double ic[19];
        int64_t ep = h->I[(size_t)I_EPISODE * h->n + e] + 1;     // as leo_reset_kernel mode 2
        h->I[(size_t)I_EPISODE * h->n + e] = ep;
        leo::sample_ic(h->P, first_env + e, ep, ic);
        for (int k = 0; k < 19; k++) { h->ics[e * 19 + k] = ic[k]; if (ics_out) ics_out[e * 19 + k] = ic[k]; }
        leo::leo_reset_env(h->P, h->S.data(), h->I.data(), h->n, e, ic, obs ? obs + 5 * e : nullptr);
    }
}
void hc_step(HC *h, const int32_t *actions, double *obs, double *reward, uint8_t *done, uint8_t *reason)
{
    double bus[leo::LEO_NM_PFIX];
    leo::MBus m; m.a = 0; m.p = bus;
    const int nch = leo_host::step_chunks(h->P);                // the same split of the interval as the CUDA kernel
    for (int64_t e = 0; e < h->n; e++)
    for (int ch = 0; ch < nch; ch++) {
        leo::StepOut o;
        const bool diag = h->P.diag && !h->force_general;
        if (h->P.grav_pfix) {  // planet-fixed degree-2 field
            if (h->P.nrw == 4) leo::leo_step_env<4, 2, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
            else if (diag) leo::leo_step_env<3, 2, true>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
            else leo::leo_step_env<3, 2, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
        } else
        if (h->P.mixed) {      // mixed-precision variant (leo_f32.cuh): the two configurations the library builds
            if (h->P.nrw == 4) leo::leo_step_env<4, true, false, true>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, h->PF, ch, nch);
            else leo::leo_step_env<3, false, true, true>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, h->PF, ch, nch);
        } else if (h->P.nrw == 4) {
            if (h->P.use_j2) leo::leo_step_env<4, true, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
            else leo::leo_step_env<4, false, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
        } else if (h->P.use_j2) {
            if (diag) leo::leo_step_env<3, true, true>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
            else leo::leo_step_env<3, true, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
        } else {
            if (diag) leo::leo_step_env<3, false, true>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
            else leo::leo_step_env<3, false, false>(h->P, h->S.data(), h->I.data(), h->n, e, m, actions[e], o, LeoParamsF(), ch, nch);
        }
        if (ch + 1 < nch) continue;
        for (int k = 0; k < 5; k++) obs[5 * e + k] = o.ob[k];
        reward[e] = o.reward; done[e] = (uint8_t)o.done; reason[e] = (uint8_t)o.reason;
    }
}
// the same two settings the C ABI offers (bskenv_set_ephemeris / bskenv_set_gravity_degree2)
void hc_set_gravity_degree2(HC *h, int enable, const double *cbar)
{
    if (!enable) h->P.grav_pfix = 0; else leo_host::set_degree2(h->P, cbar);
}
void hc_set_ephemeris(HC *h, int kind, double t0, double seg_len, int nseg, int ncoef, const double *coef)
{
    LeoEph &E = kind == 0 ? h->P.eph_sun : h->P.eph_orient;
    std::vector<double> &store = h->eph[kind];
    store.assign(coef, coef + (size_t)(nseg > 0 ? nseg : 0) * 3 * ncoef);
    E.coef = nseg > 0 ? store.data() : nullptr; E.nseg = nseg > 0 ? nseg : 0; E.ncoef = ncoef; E.t0 = t0; E.seg_len = seg_len;
}
void hc_get_state(HC *h, double *S, int64_t *I)
{
    memcpy(S, h->S.data(), h->S.size() * sizeof(double));
    memcpy(I, h->I.data(), h->I.size() * sizeof(int64_t));
}
void hc_force_general(HC *h, int on) { h->force_general = on; }
// eclipse fast path (squared cone tests + guard band) for one Sun latch time and n spacecraft positions
void hc_eclipse(HC *h, int64_t msg_ns, const double *r, int64_t n, double *out, double *sun_r)
{
    leo::SunLatch sun = leo::sun_latch(h->P, msg_ns);
    for (int64_t i = 0; i < n; i++) {
        leo::V3 p = leo::mk(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
        leo::V3 hb = sun.r - p;
        out[i] = leo::eclipse_factor(h->P, sun, p, leo::dot(p, p), hb, leo::dot(hb, hb));
    }
    sun_r[0] = sun.r.x; sun_r[1] = sun.r.y; sun_r[2] = sun.r.z;
}
int hc_dims(int *nd, int *ni) { *nd = LEO_ND; *ni = LEO_NI; return 0; }
void hc_thr_force_mapping(HC *h, const double *Lr, double *F)
{
    double f[LEO_NTHR];
    leo::thr_force_mapping(h->P, leo::mk(Lr[0], Lr[1], Lr[2]), f);
    for (int i = 0; i < LEO_NTHR; i++) F[i] = f[i];
}
}
