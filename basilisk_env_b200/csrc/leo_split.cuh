// leo_split.cuh -- the SMALL-BATCH organisation of the fused LEO decision step: TWO WARPS PER GROUP OF 32 SPACECRAFT.
//
// The same decision interval as leo_core.cuh: leo_step_env() (reference: LEOPowerAttitudeSimulator.run_sim,
// /root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py:535-644, + leoPowerAttEnv.step,
// /root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:65-145), built from the same functions, but the work
// of one spacecraft is split over two threads that sit in different warps -- i.e. on different SM sub-partitions, each
// with its own issue port and FP64 pipe -- of the same block:
//
//   dynamics thread  (split_dyn)  everything the NEXT tick's integration depends on: spacecraftPlus RK4 (DynTask, SIM:101),
//                                 wheel invariant, MRP switch, atmosphere, wheel-limit flags, Sun third body, the command
//                                 latches of the wheels and thrusters, the per-stage thruster path;
//   companion thread (split_env)  everything that only OBSERVES the state: EnvTask (eclipse, solar panel, battery:
//                                 SIM:102-103, 311-345) one tick behind, and the flight-software pass (SIM:383-386,
//                                 hillPoint / attTrackingError / MRP_Feedback / rwMotorTorque / desat chain) -- its wheel
//                                 command is only latched AFTER the integration of the tick it runs in, so it has a whole
//                                 tick to get there.
//
// Why: with 4096 envs (BASELINE configs[1]) there are 128 warps for 592 sub-partitions; one warp per group runs its 1800
// serial ticks alone at its dependent-issue latency (3.3 ms per interval).  Measured by compiling the two companion
// blocks out of the one-thread kernel: 3.78 ms -> 1.97 ms; they are 48 % of the chain and none of it feeds the next
// tick.  (Splitting one RK4 stage over lanes or warps does not pay, and a THIRD warp that takes the flight software off
// the companion -- control half handed over first, guidance half at leisure -- was built and measured 4 % slower than
// this form, 2.50 against 2.39 ms at 4096 envs: what it saves in waiting it pays in barrier operations.  DESIGN.md 5b.)
//
// Hand-off: a per-lane mailbox in shared memory, double-buffered by tick parity, and two named barriers per group
// (PTX barrier.sync / barrier.arrive with 64 threads): TICK, by both warps once per tick (the state after tick j is in
// slot j & 1; the companion has consumed slot (j - 1) & 1), and FSW, arrive by the companion / sync by the dynamics
// warp in the ticks whose flight-software outputs it is about to latch.  Ownership of the persistent state is
// disjoint: the companion owns the desat chain's fields (I_INITREQ .. F_THRCMD) and the bus fields M_GUID .. M_RWCMD;
// the dynamics thread owns everything else; F_THRCMD and M_RWCMD cross at the FSW barrier.
//
// Results: identical arithmetic, operation by operation, to leo_step_env (same functions on the same operands); the
// parity tests compare both organisations with the oracle and with each other.
#pragma once
#include "leo_core.cuh"

#if defined(__CUDACC__)
namespace leo {

enum SplitField : int {
    DX_R = 0, DX_V = 3, DX_S = 6, DX_W = 9, DX_WHL = 12, DX_R2 = 16, DX_IR = 17, DX_H = 18, SPLIT_SLOT = 19,   // per tick, two slots
    DB_RAN = 2 * SPLIT_SLOT,  // companion -> dynamics: return value of fsw_pass
    DB_QUIET,                 // dynamics -> companion: desat chain confirmed quiet by the thruster latch
    DB_CHARGE, DB_SHADOW,     // companion -> dynamics at the end of the call
    SPLIT_NF
};
#define LEO_SPLIT_BOX_BYTES ((size_t)leo::SPLIT_NF * LEO_BLOCK * sizeof(double))

__device__ __forceinline__ void split_sync(int id) { asm volatile("barrier.sync %0, 64;" : : "r"(id) : "memory"); }
__device__ __forceinline__ void split_arrive(int id)
{
    __threadfence_block();
    asm volatile("barrier.arrive %0, 64;" : : "r"(id) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// dynamics thread
// ------------------------------------------------------------------------------------------------------------------
template <int NRW, int J2, bool DIAG>
__device__ __forceinline__ void split_dyn(const LeoParams &P, double *__restrict__ S, int64_t *__restrict__ I, int64_t stride, int64_t e,
                                        MBus m, MBus box, int bar, int action, StepOut &out, int chunk, int n_chunks)
{
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
    // ---------------- load (as leo_step_env) ----------------
    Dyn x;
    StageIn a;
    x.r = mk(SD(F_R), SD(F_R + 1), SD(F_R + 2));
    x.v = mk(SD(F_V), SD(F_V + 1), SD(F_V + 2));
    x.s = mk(SD(F_SIG), SD(F_SIG + 1), SD(F_SIG + 2));
    x.w = mk(SD(F_OMG), SD(F_OMG + 1), SD(F_OMG + 2));
    double C[NRW], uJ[NRW];
    a.tau_u = mk(0., 0., 0.); a.HB = mk(0., 0., 0.);
#pragma unroll
    for (int i = 0; i < NRW; i++) {
        const double u = SD(F_UCUR + i);
        mst(m, M_U + i, u);
        uJ[i] = u * P.invJs[i];
        C[i] = SD(F_WHL + i) + dot(arr(P.gs[i]), x.w);
        a.tau_u = a.tau_u + arr(P.gs[i]) * u;
        a.HB = a.HB + arr(P.gs[i]) * (P.Js[i] * C[i]);
    }
    for (int f = 0; f < LEO_M_MIRROR; f++) mst(m, f, SD(F_GUID + f));
    a.rho = SD(F_RHO);
    {
        const V3 L_ext = mk(SD(F_LDIST), SD(F_LDIST + 1), SD(F_LDIST + 2));
        mst3(m, M_LEXT, L_ext); mst3(m, M_FM, mk(0., 0., 0.));
        a.Lc = L_ext - a.tau_u;
    }
    mst3(m, M_LTHR, mk(0., 0., 0.));
    const int64_t tick = SI(I_TICK);
    int mask = (int)SI(I_MASK);
    int thr_factor = (int)SI(I_THRFACTOR), thr_active = (int)SI(I_THRACTIVE), rw_sat = (int)SI(I_RWSAT);
    int nswitch = 0;
    // mode switch (SIM:543-588): the task mask; the Reset calls of action 2 touch the desat chain, which the companion owns
    if (chunk > 0) {}
    else if (action == 0) mask = LEO_TASK_NADIR | LEO_TASK_MRP;
    else if (action == 1) mask = LEO_TASK_SUN | LEO_TASK_MRP;
    else if (action == 2) mask = LEO_TASK_SUN | LEO_TASK_MRP | LEO_TASK_DESAT;
    mst(m, M_TNEXT, -1.0);

    const bool first = tick < 0;
    const int tpf = P.ticks_per_fsw;
    const int ticks_step = tpf * P.fsw_per_step;
    const int ticks = ticks_step / n_chunks;
    const int64_t n_base = (first ? 0 : tick) + 1;
    const int64_t n_step0 = ((n_base - 1) / ticks_step) * ticks_step;
    const int64_t n_end = n_step0 + ticks_step;
    const double dyn_d = (double)P.dyn_ns;
    double sun_d = (double)(n_step0 * P.dyn_ns);
    sun_latch_to_bus(P, m, n_step0 * P.dyn_ns);
    if (J2 == 2) pfix_latch_to_bus(P, m, n_step0 * P.dyn_ns);
    a.dtp = 0.;
    const int j0 = first ? -1 : 0;
    const int jlo = __any_sync(0xffffffffu, first) ? -1 : 0;   // warp-uniform loop start: every lane takes every barrier
    int phase = (int)((n_base + j0) % tpf);
    // ticks >= 0 of all lanes share one flight-software phase unless a foreign tick count was injected (set_state): no vote then
    const int ph0 = (int)((n_base) % tpf);
    const bool ph_uniform = __all_sync(0xffffffffu, ph0 == __shfl_sync(0xffffffffu, ph0, 0));
    double now_d = (double)((n_base + j0) * P.dyn_ns);
    int desat_ran = 0;
    double newTime = t_mul(now_d, LEO_NANO2SEC);
    double prevTime = j0 < 0 ? 0.0 : t_mul(now_d - dyn_d, 1e-9);
    double h = t_sub(newTime, prevTime);
    double dtsm = t_mul((j0 < 0 ? 0.0 : now_d - dyn_d) - sun_d, LEO_NANO2SEC) + 0.5 * h;
    const double dyn_s = t_mul(dyn_d, LEO_NANO2SEC);
    a.gsun = sun_window(P, m, x, dtsm - 0.5 * h, h, tpf - phase, dyn_s);        // held per flight-software period, as leo_step_env

    // the state before the first tick, for a flight-software pass that is due in it
    {
        double W[NRW];
        wheel_speeds<NRW, DIAG>(P, a.HB, C, x.w, W);
        for (int sl = 0; sl <= SPLIT_SLOT; sl += SPLIT_SLOT) {        // both slots: a lane that sits out tick -1 publishes nothing in it
            mst3(box, sl + DX_R, x.r); mst3(box, sl + DX_V, x.v); mst3(box, sl + DX_S, x.s); mst3(box, sl + DX_W, x.w);
#pragma unroll
            for (int i = 0; i < NRW; i++) mst(box, sl + DX_WHL + i, W[i]);
        }
        mst(box, DB_QUIET, 0.0);
    }
    __syncwarp();
    split_sync(bar);                                           // START: bus mirror, Sun latch and mailbox are in place

#pragma unroll 1
    for (int j = jlo; j < ticks; j++) {
        const bool on = j >= j0;
        const bool fsw_now = on && phase == 0;
        const bool fsw_any = (ph_uniform && j >= 0) ? fsw_now : __any_sync(0xffffffffu, fsw_now);
        double W[NRW];
        int lim = 0;
        if (on) {
            bool wrapped = false;
            if (LEO_RARE(fsw_now)) {
                const int64_t n = n_base + j;
                rw_sat |= 2;
                if (n > 0 && n == n_end) {
                    sun_d = now_d; sun_latch_to_bus(P, m, n * P.dyn_ns); wrapped = true;
                    if (J2 == 2) pfix_latch_to_bus(P, m, n * P.dyn_ns);
                }
                // Sun third body of the ticks up to the next pass (as leo_step_env)
                a.gsun = sun_window(P, m, x, t_mul((j < 0 ? 0.0 : now_d - dyn_d) - sun_d, LEO_NANO2SEC), h, tpf, dyn_s);
            }
            // ================= DynTask =================
            a.h = h;
            if (LEO_RARE(wrapped || (thr_active && !(newTime + LEO_THR_MARGIN <= mld(m, M_TNEXT))))) {
                const double prev_d = j < 0 ? 0.0 : now_d - dyn_d;
                const double tBefore = t_sub(newTime, h);
                SunDt dts;
                dts.d0 = t_mul(prev_d - sun_d, 1e-9); dts.dm = dts.d0 + 0.5 * h; dts.d1 = dts.d0 + h;
                if (wrapped) dts = sun_dt_wrapped(prev_d, sun_d, tBefore, prevTime, h);
                double tauPrev = 0.0;
                if (j >= 0) {
                    const double ppT = (n_base + j > 1) ? t_mul(prev_d - dyn_d, 1e-9) : 0.0;
                    const double ph = t_sub(prevTime, ppT);
                    tauPrev = t_add(t_sub(prevTime, ph), ph);
                }
                ThrEventOut o = rk4_general<J2, DIAG>(P, S, stride, e, m, x, a, dts, tBefore, tauPrev, thr_factor, thr_active);
                x = o.x; thr_factor = o.factor; thr_active = o.active;
                ThrRefresh th = thr_refresh(P, S, stride, e, m, thr_active ? thr_factor : 0, a.tau_u);
                a.Lc = th.Lc; mst3(m, M_LTHR, th.L_thr); mst(m, M_TNEXT, th.t_next);
            } else {
                if (J2 == 2) a.dtp = dtsm;
                x = rk4_step<J2, DIAG>(P, x, a, thr_active != 0, m);
            }
            // ================= what the next tick depends on =================
            a.HB = a.HB + a.tau_u * h;
            if (!DIAG) {
#pragma unroll
                for (int i = 0; i < NRW; i++) C[i] = fmad(uJ[i], h, C[i]);
            }
            {   // MRP shadow-set switch: a branch here (the reciprocal of the select form sits on the lone warp's chain)
                const double s2 = dot(x.s, x.s);
                if (LEO_RARE(s2 > 1.0000000000000002)) {
                    const double f = -frcp(s2);
                    x.s = mk(x.s.x * f, x.s.y * f, x.s.z * f);
                    nswitch++;
                }
            }
            const double r2 = dot(x.r, x.r), ir = rsq(r2);
            a.rho = P.rho0 * exp_bounded(-(r2 * ir - P.Rp_atmo) * P.inv_H);
            wheel_speeds<NRW, DIAG>(P, a.HB, C, x.w, W);
#pragma unroll
            for (int i = 0; i < NRW; i++) lim |= (fabs(W[i]) >= P.Om_max[i] && P.Om_max[i] > 0.0) ? 1 : 0;
            // publish the state after this tick
            {
                const int sl = (j & 1) * SPLIT_SLOT;
                mst3(box, sl + DX_R, x.r); mst3(box, sl + DX_V, x.v); mst3(box, sl + DX_S, x.s); mst3(box, sl + DX_W, x.w);
#pragma unroll
                for (int i = 0; i < NRW; i++) mst(box, sl + DX_WHL + i, W[i]);
                mst(box, sl + DX_R2, r2); mst(box, sl + DX_IR, ir); mst(box, sl + DX_H, h);
            }
            now_d += dyn_d;
            phase = (phase + 1 == tpf) ? 0 : phase + 1;
            prevTime = newTime;
            newTime = t_mul(now_d, LEO_NANO2SEC);
            h = t_sub(newTime, prevTime);
            dtsm = t_mul((now_d - dyn_d) - sun_d, LEO_NANO2SEC) + 0.5 * h;
        }
        // ================= flight-software outputs of this tick (written by the companion while we integrated) =================
        if (fsw_any) {
            __syncwarp();
            split_sync(bar + 1);
            if (fsw_now) desat_ran = (int)mld(box, DB_RAN);
        }
        if (on && LEO_RARE(rw_sat | lim | desat_ran)) {
            // (out of line even for the plain wheel latch of every pass: inlining it here measured 2 % slower)
            PostOut<NRW> po = post_tick_events<NRW>(P, S, I, stride, e, m, W, mld3(m, M_LTHR), desat_ran, (int64_t)(now_d - dyn_d), thr_factor);
#pragma unroll
            for (int i = 0; i < NRW; i++) uJ[i] = po.uJ[i];
            a.Lc = po.Lc; a.tau_u = po.tau_u;
            rw_sat = lim;
            if (po.thr_active >= 0) { thr_active = po.thr_active; mst(m, M_TNEXT, -1.0); }
            if (desat_ran) mst(box, DB_QUIET, (double)po.quiet);
            desat_ran = 0;
        }
        __syncwarp();
        split_sync(bar);                                     // TICK
    }
    double W[NRW];
    wheel_speeds<NRW, DIAG>(P, a.HB, C, x.w, W);
    split_sync(bar);                                           // FINAL: battery charge and shadow factor of the last tick
    const double charge = mld(box, DB_CHARGE), shadow = mld(box, DB_SHADOW);
    for (int f = 0; f < LEO_M_MIRROR; f++) SD(F_GUID + f) = mld(m, f);
    if (chunk + 1 < n_chunks) {
        SD(F_R) = x.r.x; SD(F_R + 1) = x.r.y; SD(F_R + 2) = x.r.z;
        SD(F_V) = x.v.x; SD(F_V + 1) = x.v.y; SD(F_V + 2) = x.v.z;
        SD(F_SIG) = x.s.x; SD(F_SIG + 1) = x.s.y; SD(F_SIG + 2) = x.s.z;
        SD(F_OMG) = x.w.x; SD(F_OMG + 1) = x.w.y; SD(F_OMG + 2) = x.w.z;
#pragma unroll
        for (int i = 0; i < NRW; i++) { SD(F_WHL + i) = W[i]; SD(F_UCUR + i) = mld(m, M_U + i); }
        SD(F_RHO) = a.rho; SD(F_E) = charge; SD(F_SHADOW) = shadow;
        SI(I_TICK) = n_base - 1 + ticks; SI(I_MASK) = mask; SI(I_SWITCH) = SI(I_SWITCH) + nswitch;
        SI(I_THRFACTOR) = thr_factor; SI(I_THRACTIVE) = thr_active; SI(I_RWSAT) = rw_sat;
        out.done = 0; out.reason = 0; out.reward = 0.;
        __syncwarp();
        split_sync(bar);                                         // END: the state is in memory for the next chunk's companion
        return;
    }
    // ---------------- observation sampling (SIM:598-642) + gym bookkeeping (ENV:98-145) ----------------
    double ob0 = norm(mld3(m, M_GUID));
    double ob1 = norm(x.w);
    double wn = 0.;
#pragma unroll
    for (int i = 0; i < NRW; i++) wn += W[i] * W[i];
    const double E = charge;
    double ob2 = sqrt(wn), ob3 = E / 3600., ob4 = shadow;
    int sim_over = norm(x.r) < P.decay_radius;
    int64_t curr_step = SI(I_STEP);
    int over = (int)SI(I_OVER), reason = 0;
    if (curr_step >= P.max_length) { over = 1; reason |= 1; }
    double reward = 0.;
    if (action == 0) reward = fabs(P.reward_mult / (1. + ob0 * ob0));
    double ret = SD(F_EPRET) + reward;
    SD(F_OBS) = ob0; SD(F_OBS + 1) = ob1; SD(F_OBS + 2) = ob2; SD(F_OBS + 3) = ob3; SD(F_OBS + 4) = ob4;
    ob2 = ob2 / P.wheel_limit;
    ob3 = ob3 / P.power_max;
    if (ob2 > 1.) { over = 1; reward -= P.failure_penalty; ret -= P.failure_penalty; reason |= 2; }
    if (ob3 == 0.) { over = 1; reward -= P.failure_penalty; ret -= P.failure_penalty; reason |= 4; }
    if (sim_over) { over = 1; reason |= 8; }
    out.ob[0] = ob0; out.ob[1] = ob1; out.ob[2] = ob2; out.ob[3] = ob3; out.ob[4] = ob4;
    out.reward = reward; out.done = over; out.reason = reason;

    SD(F_R) = x.r.x; SD(F_R + 1) = x.r.y; SD(F_R + 2) = x.r.z;
    SD(F_V) = x.v.x; SD(F_V + 1) = x.v.y; SD(F_V + 2) = x.v.z;
    SD(F_SIG) = x.s.x; SD(F_SIG + 1) = x.s.y; SD(F_SIG + 2) = x.s.z;
    SD(F_OMG) = x.w.x; SD(F_OMG + 1) = x.w.y; SD(F_OMG + 2) = x.w.z;
#pragma unroll
    for (int i = 0; i < NRW; i++) { SD(F_WHL + i) = W[i]; SD(F_UCUR + i) = mld(m, M_U + i); }
    SD(F_RHO) = a.rho; SD(F_E) = E; SD(F_SHADOW) = shadow; SD(F_EPRET) = ret;
    SI(I_TICK) = n_end; SI(I_STEP) = curr_step + 1; SI(I_MASK) = mask; SI(I_SWITCH) = SI(I_SWITCH) + nswitch;
    SI(I_THRFACTOR) = thr_factor; SI(I_THRACTIVE) = thr_active; SI(I_OVER) = over; SI(I_RWSAT) = rw_sat;
    __syncwarp();
    split_sync(bar);                                           // END
#undef SD
#undef SI
}

// ------------------------------------------------------------------------------------------------------------------
// companion thread: flight software + EnvTask
// ------------------------------------------------------------------------------------------------------------------
template <int NRW>
__device__ __forceinline__ void split_env(const LeoParams &P, double *__restrict__ S, int64_t *__restrict__ I, int64_t stride, int64_t e,
                                        MBus m, MBus box, int bar, int action, int chunk, int n_chunks)
{
#define SD(f) S[(int64_t)(f) * stride + e]
#define SI(f) I[(int64_t)(f) * stride + e]
    const int64_t tick = SI(I_TICK);
    int mask = (int)SI(I_MASK);
    double charge = SD(F_E), shadow = SD(F_SHADOW);
    if (chunk > 0) {}
    else if (action == 0) mask = LEO_TASK_NADIR | LEO_TASK_MRP;
    else if (action == 1) mask = LEO_TASK_SUN | LEO_TASK_MRP;
    else if (action == 2) {
        mask = LEO_TASK_SUN | LEO_TASK_MRP | LEO_TASK_DESAT;
        SI(I_INITREQ) = 1;                                   // thrDesatControlWrap.Reset (SIM:580)
        SI(I_DUMPPRIOR) = 0; SI(I_DUMPCNT) = 0; SI(I_LASTDH) = 0;   // thrDumpWrap.Reset (SIM:581)
        for (int k = 0; k < LEO_NTHR; k++) SD(F_THRREM + k) = 0.0;
    }
    const bool first = tick < 0;
    const int tpf = P.ticks_per_fsw;
    const int ticks_step = tpf * P.fsw_per_step;
    const int ticks = ticks_step / n_chunks;
    const int64_t n_base = (first ? 0 : tick) + 1;
    const int64_t n_step0 = ((n_base - 1) / ticks_step) * ticks_step;
    const int64_t n_end = n_step0 + ticks_step;
    const int64_t sun_ns = n_step0 * P.dyn_ns;               // every pass of this call sees the interval's first Sun message
    const int j0 = first ? -1 : 0;
    const int jlo = __any_sync(0xffffffffu, first) ? -1 : 0;
    const int ph0 = (int)((n_base) % tpf);
    const bool ph_uniform = __all_sync(0xffffffffu, ph0 == __shfl_sync(0xffffffffu, ph0, 0));

    split_sync(bar);                                           // START
    V3 sun_r = mld3(m, M_SUNR);
    double ec[6] = {mld(m, M_ECL), mld(m, M_ECL + 1), mld(m, M_ECL + 2), mld(m, M_ECL + 3), mld(m, M_ECL + 4), mld(m, M_ECL + 5)};

    // flight-software pass of tick jj, from the state after tick jj - 1 (slot (jj - 1) & 1)
    auto fsw_if_due = [&](int jj) {
        const bool due = jj >= j0 && (int)((n_base + jj) % tpf) == 0;
        if (!((ph_uniform && jj >= 0) ? due : __any_sync(0xffffffffu, due))) return;
        if (due) {
            const int64_t n = n_base + jj;
            const int sl = ((jj - 1) & 1) * SPLIT_SLOT;
            Dyn x;
            x.r = mld3(box, sl + DX_R); x.v = mld3(box, sl + DX_V); x.s = mld3(box, sl + DX_S); x.w = mld3(box, sl + DX_W);
            double W[NRW];
#pragma unroll
            for (int i = 0; i < NRW; i++) W[i] = mld(box, sl + DX_WHL + i);
            const int quiet = (int)mld(box, DB_QUIET);
            const int ran = fsw_pass<NRW>(P, S, I, stride, e, m, mask, n, n * P.dyn_ns, x, W, sun_ns, quiet);
            mst(box, DB_RAN, (double)ran);
        }
        __syncwarp();
        split_arrive(bar + 1);
    };
    fsw_if_due(jlo);

#pragma unroll 1
    for (int j = jlo; j < ticks; j++) {
        split_sync(bar);                                     // TICK j: slot j & 1 holds the state after tick j
        if (j + 1 < ticks) fsw_if_due(j + 1);                  // first: the dynamics warp waits for this one
        if (j < j0) continue;
        const int64_t n = n_base + j;
        if (n > 0 && n == n_end) {                             // the tick ran under the NEXT interval's Sun message (quirk Q18)
            sun_r = mld3(m, M_SUNR);
#pragma unroll
            for (int q = 0; q < 6; q++) ec[q] = mld(m, M_ECL + q);
        }
        // ================= EnvTask: eclipse cone tests, solar-panel geometry, battery =================
        const int sl = (j & 1) * SPLIT_SLOT;
        const V3 xr = mld3(box, sl + DX_R), xs = mld3(box, sl + DX_S);
        const double r2 = mld(box, sl + DX_R2), ir = mld(box, sl + DX_IR), h_this = mld(box, sl + DX_H);
        const V3 r_SB = sun_r - xr;
        const double d2 = dot(r_SB, r_SB);
        const double id = rsq(d2);
        bool penumbra;
        shadow = eclipse_cones(P, ec, sun_r, xr, r2, d2, penumbra);
        const double rdh = dot(xr, r_SB);
        MrpRot R = mrp_rot(xs);
        V3 n_N = rot_NB(R, xs, arr(P.nHat_B));
        double proj = dot(n_N, r_SB) * id;
        if (proj < 0.) proj = 0.;
        const double pgeo = P.panel_coef * proj * (id * id);
        if (LEO_RARE(penumbra)) {
#ifdef LEO_LITERAL_ECLIPSE
            shadow = penumbra_literal(P, sun_r, xr);
#else
            shadow = penumbra_fraction(P, ir, id, rdh);
#endif
        }
        {
            const double panel = pgeo * shadow;
            double E = charge + (panel + P.sink_power) * h_this;
            if (E > P.capacity) E = P.capacity;
            if (E < 0.) E = 0.;
            charge = j >= 0 ? E : charge;
        }
    }
    mst(box, DB_CHARGE, charge); mst(box, DB_SHADOW, shadow);
    __syncwarp();
    split_sync(bar);                                           // FINAL
    split_sync(bar);                                           // END
#undef SD
#undef SI
}

}  // namespace leo
#endif  // __CUDACC__
