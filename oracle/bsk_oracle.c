/*
 * bsk_oracle.c -- scalar FP64 CPU restatement of the LEO power/attitude hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see bsk_oracle.h).  PARITY UNPINNED: Basilisk is absent from
 * /root/reference and from this image, and the reference holds no golden outputs.
 *
 * Structure deliberately mirrors Basilisk 1.x rather than the fused CUDA kernel, so that the two
 * are independent restatements of the same specification:
 *   - an integer-nanosecond priority scheduler (sim_model / sys_process / sys_model_task),
 *   - a message bus of structs with {write time, update counter} headers,
 *   - one struct + Update function per Basilisk module,
 *   - a message logger sampled at the decision period.
 * Wiring, constants and ordering follow
 *   /root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py           (cited as SIM:line)
 *   /root/reference/basilisk_env/simulators/dynamics/effectorPrimatives/actuatorPrimatives.py (AP:line)
 *   /root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py               (ENV:line)
 * Module math follows the published Basilisk 1.x algorithms ([BSK: path] = upstream AVS-Lab path,
 * not present here; see SURVEY.md Appendix A and docs/PHYSICS_SPEC.md).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -fopenmp (oracle/Makefile).
 */
#include "bsk_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NANO2SEC 1e-9
#define DB0_EPS 1e-30
#define MAX_THR 8
#define MAX_RW 4

/* ---- constants ([BSK: src/utilities/astroConstants.h, macros.py, simIncludeGravBody.py]) ---- */
static const double MU_EARTH = 0.3986004415e15;      /* createEarth().mu, also leo_orbit.py:30 */
static const double MU_SUN = 1.32712440018e20;       /* createSun().mu */
static const double REQ_EARTH_KM = 6378.1366;        /* orbitalMotion.REQ_EARTH / astroConstants REQ_EARTH */
static const double REQ_SUN_KM = 695000.0;           /* astroConstants REQ_SUN (confidence L) */
static const double AU_KM = 149597870.693;           /* astroConstants AU */
static const double SOLAR_FLUX_EARTH = 1372.5398;    /* astroConstants SOLAR_FLUX_EARTH */
static const double J2_EARTH = 1.08262668355e-3;     /* GGM03S degree-2 zonal, only when use_j2 */
static const double RPM = 0.10471975511965977;       /* macros.RPM = 2*pi/60 */
#define PI_D 3.14159265358979323846

static uint64_t sec2nano(double x) { return (uint64_t)(x * 1e9 + 0.5); } /* macros.sec2nano */

/* ------------------------------ linear algebra (linearAlgebra.c style) ------------------------- */
static double v3Dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3Norm(const double a[3]) { return sqrt(v3Dot(a, a)); }
static void v3Copy(const double a[3], double r[3]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static void v3SetZero(double r[3]) { r[0] = r[1] = r[2] = 0.0; }
static void v3Scale(double s, const double a[3], double r[3]) { r[0] = s * a[0]; r[1] = s * a[1]; r[2] = s * a[2]; }
static void v3Add(const double a[3], const double b[3], double r[3]) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static void v3Subtract(const double a[3], const double b[3], double r[3]) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void v3Cross(const double a[3], const double b[3], double r[3])
{
    double t[3];
    t[0] = a[1] * b[2] - a[2] * b[1];
    t[1] = a[2] * b[0] - a[0] * b[2];
    t[2] = a[0] * b[1] - a[1] * b[0];
    v3Copy(t, r);
}
static void v3Normalize(const double a[3], double r[3])
{
    double n = v3Norm(a);
    if (n > DB0_EPS) v3Scale(1. / n, a, r); else v3SetZero(r);
}
static void m33MultV3(double m[3][3], const double v[3], double r[3])
{
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = m[i][0] * v[0] + m[i][1] * v[1] + m[i][2] * v[2];
    v3Copy(t, r);
}
static void m33tMultV3(double m[3][3], const double v[3], double r[3])
{
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = m[0][i] * v[0] + m[1][i] * v[1] + m[2][i] * v[2];
    v3Copy(t, r);
}
static double m33Determinant(double m[3][3])
{
    return m[0][0] * m[1][1] * m[2][2] + m[0][1] * m[1][2] * m[2][0] + m[0][2] * m[1][0] * m[2][1]
         - m[0][0] * m[1][2] * m[2][1] - m[0][1] * m[1][0] * m[2][2] - m[0][2] * m[1][1] * m[2][0];
}
static void m33Inverse(double m[3][3], double r[3][3])
{ /* adjugate / determinant, as linearAlgebra.c m33Inverse and Eigen's 3x3 inverse */
    double det = m33Determinant(m), t[3][3];
    t[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det;
    t[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det;
    t[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det;
    t[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
    t[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det;
    t[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
    t[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det;
    t[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det;
    t[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
    memcpy(r, t, sizeof(t));
}
static double safeAsin(double x) { return x > 1. ? asin(1.) : (x < -1. ? asin(-1.) : asin(x)); }
static double safeAcos(double x) { return x > 1. ? acos(1.) : (x < -1. ? acos(-1.) : acos(x)); }

/* ------------------------------ rigid body kinematics ([BSK: RigidBodyKinematics.c]) ----------- */
void orc_MRP2C(const double q[3], double C[3][3])
{ /* [BN] from sigma_BN */
    double q1 = q[0], q2 = q[1], q3 = q[2];
    double d1 = v3Dot(q, q);
    double S = 1 - d1;
    double d = (1 + d1) * (1 + d1);
    C[0][0] = 4 * (2 * q1 * q1 - d1) + S * S;
    C[0][1] = 8 * q1 * q2 + 4 * q3 * S;
    C[0][2] = 8 * q1 * q3 - 4 * q2 * S;
    C[1][0] = 8 * q2 * q1 - 4 * q3 * S;
    C[1][1] = 4 * (2 * q2 * q2 - d1) + S * S;
    C[1][2] = 8 * q2 * q3 + 4 * q1 * S;
    C[2][0] = 8 * q3 * q1 + 4 * q2 * S;
    C[2][1] = 8 * q3 * q2 - 4 * q1 * S;
    C[2][2] = 4 * (2 * q3 * q3 - d1) + S * S;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = (1. / d) * C[i][j];
}
static void C2EP(double C[3][3], double b[4])
{ /* Sheppard's method */
    double tr = C[0][0] + C[1][1] + C[2][2], b2[4], max;
    int i = 0;
    b2[0] = (1 + tr) / 4.;
    b2[1] = (1 + 2 * C[0][0] - tr) / 4.;
    b2[2] = (1 + 2 * C[1][1] - tr) / 4.;
    b2[3] = (1 + 2 * C[2][2] - tr) / 4.;
    max = b2[0];
    for (int j = 1; j < 4; j++) if (b2[j] > max) { i = j; max = b2[j]; }
    switch (i) {
    case 0:
        b[0] = sqrt(b2[0]);
        b[1] = (C[1][2] - C[2][1]) / 4 / b[0];
        b[2] = (C[2][0] - C[0][2]) / 4 / b[0];
        b[3] = (C[0][1] - C[1][0]) / 4 / b[0];
        break;
    case 1:
        b[1] = sqrt(b2[1]);
        b[0] = (C[1][2] - C[2][1]) / 4 / b[1];
        if (b[0] < 0) { b[1] = -b[1]; b[0] = -b[0]; }
        b[2] = (C[0][1] + C[1][0]) / 4 / b[1];
        b[3] = (C[2][0] + C[0][2]) / 4 / b[1];
        break;
    case 2:
        b[2] = sqrt(b2[2]);
        b[0] = (C[2][0] - C[0][2]) / 4 / b[2];
        if (b[0] < 0) { b[2] = -b[2]; b[0] = -b[0]; }
        b[1] = (C[0][1] + C[1][0]) / 4 / b[2];
        b[3] = (C[1][2] + C[2][1]) / 4 / b[2];
        break;
    default:
        b[3] = sqrt(b2[3]);
        b[0] = (C[0][1] - C[1][0]) / 4 / b[3];
        if (b[0] < 0) { b[3] = -b[3]; b[0] = -b[0]; }
        b[1] = (C[2][0] + C[0][2]) / 4 / b[3];
        b[2] = (C[1][2] + C[2][1]) / 4 / b[3];
        break;
    }
}
void orc_C2MRP(double C[3][3], double q[3])
{
    double b[4] = {1., 0., 0., 0.};
    C2EP(C, b);
    q[0] = b[1] / (1 + b[0]);
    q[1] = b[2] / (1 + b[0]);
    q[2] = b[3] / (1 + b[0]);
}
void orc_addMRP(const double q1[3], const double q2[3], double result[3])
{ /* singularity guard + inner-set mapping form (confidence M) */
    double v1[3], v2[3], s1[3], det, mag, res[3];
    v3Copy(q1, s1);
    det = (1 + v3Dot(s1, s1) * v3Dot(q2, q2) - 2 * v3Dot(s1, q2));
    if (fabs(det) < 0.1) {
        mag = v3Dot(s1, s1);
        v3Scale(-1. / mag, s1, s1);
        det = (1 + v3Dot(s1, s1) * v3Dot(q2, q2) - 2 * v3Dot(s1, q2));
    }
    v3Cross(s1, q2, v1);
    v3Scale(2., v1, v1);
    v3Scale(1 - v3Dot(q2, q2), s1, res);
    v3Scale(1 - v3Dot(s1, s1), q2, v2);
    v3Add(res, v2, res);
    v3Add(res, v1, res);               /* [FN(Q)] = [FB(q2)][BN(q1)]: +2 q1 x q2 (Schaub & Junkins eq. 3.155) */
    v3Scale(1 / det, res, res);
    mag = v3Dot(res, res);
    if (mag > 1.0) v3Scale(-1. / mag, res, res);
    v3Copy(res, result);
}
void orc_subMRP(const double q1[3], const double q2[3], double result[3])
{
    double v1[3], v2[3], s1[3], det, mag, res[3];
    v3Copy(q1, s1);
    det = (1. + v3Dot(s1, s1) * v3Dot(q2, q2) + 2. * v3Dot(s1, q2));
    if (fabs(det) < 0.1) {
        mag = v3Dot(s1, s1);
        v3Scale(-1.0 / mag, s1, s1);
        det = (1. + v3Dot(s1, s1) * v3Dot(q2, q2) + 2. * v3Dot(s1, q2));
    }
    v3Cross(s1, q2, v1);
    v3Scale(2., v1, v1);
    v3Scale(1. - v3Dot(q2, q2), s1, res);
    v3Scale(1. - v3Dot(s1, s1), q2, v2);
    v3Subtract(res, v2, res);
    v3Add(res, v1, res);
    v3Scale(1. / det, res, res);
    mag = v3Dot(res, res);
    if (mag > 1.0) v3Scale(-1. / mag, res, res);
    v3Copy(res, result);
}

/* ------------------------------ IC helpers ([BSK: orbitalMotion.py elem2rv]) -------------------- */
void orc_elem2rv(double mu, double a, double e, double i, double Omega, double omega, double f,
                 double rVec[3], double vVec[3])
{ /* non-rectilinear branch; leo_orbit.py:38 always lands here (a>0, e<1) */
    double p = a * (1.0 - e * e);
    double r = p / (1.0 + e * cos(f));
    double theta = omega + f;
    rVec[0] = r * (cos(theta) * cos(Omega) - cos(i) * sin(theta) * sin(Omega));
    rVec[1] = r * (cos(theta) * sin(Omega) + cos(i) * sin(theta) * cos(Omega));
    rVec[2] = r * (sin(theta) * sin(i));
    double h = sqrt(mu * p);
    vVec[0] = -mu / h * (cos(Omega) * (e * sin(omega) + sin(theta)) + cos(i) * (e * cos(omega) + cos(theta)) * sin(Omega));
    vVec[1] = -mu / h * (sin(Omega) * (e * sin(omega) + sin(theta)) - cos(i) * (e * cos(omega) + cos(theta)) * cos(Omega));
    vVec[2] = mu / h * (e * cos(omega) + cos(theta)) * sin(i);
}

/* ------------------------------ Sun ephemeris (substitute for SPICE de430) --- */
/* Epoch '2021 MAY 04 07:47:48.965 (UTC)' (SIM:219).  Astronomical-Almanac low-precision solar
 * coordinates, Earth-centred, mean equator/equinox treated as J2000.  Velocity = analytic derivative.
 * Documented deviation from a SPICE-driven run (~1e-4 relative). */
#define EPOCH_DAYS_TT_FROM_J2000 (7793.5 + (28068.965 + 69.184) / 86400.0)
void orc_sun_ephemeris(double t, double r[3], double v[3], double *j2000_et)
{
    const double D2R = PI_D / 180.0;
    double n = EPOCH_DAYS_TT_FROM_J2000 + t / 86400.0;
    double nd = 1.0 / 86400.0;
    double L = (280.460 + 0.9856474 * n) * D2R, Ld = 0.9856474 * D2R * nd;
    double g = (357.528 + 0.9856003 * n) * D2R, gd = 0.9856003 * D2R * nd;
    double sg = sin(g), cg = cos(g), s2g = sin(2 * g), c2g = cos(2 * g);
    double lam = L + (1.915 * sg + 0.020 * s2g) * D2R;
    double lamd = Ld + (1.915 * cg + 0.040 * c2g) * D2R * gd;
    double eps = (23.439 - 4e-7 * n) * D2R, epsd = -4e-7 * D2R * nd;
    double AUm = AU_KM * 1000.0;
    double R = (1.00014 - 0.01671 * cg - 0.00014 * c2g) * AUm;
    double Rd = (0.01671 * sg + 0.00028 * s2g) * gd * AUm;
    double sl = sin(lam), cl = cos(lam), se = sin(eps), ce = cos(eps);
    double u[3] = {cl, ce * sl, se * sl};
    double ud[3] = {-sl * lamd, ce * cl * lamd - se * sl * epsd, se * cl * lamd + ce * sl * epsd};
    for (int k = 0; k < 3; k++) { r[k] = R * u[k]; v[k] = Rd * u[k] + R * ud[k]; }
    if (j2000_et) *j2000_et = EPOCH_DAYS_TT_FROM_J2000 * 86400.0 + t;
}

/* Sun relative to the Mars barycentre, J2000 equatorial axes: analytic stand-in for SPICE de430 with
 * zeroBase = "mars barycenter" (opNav_models/BSK_OpNavDynamics.py:396-401).  Keplerian elements of Mars and their
 * rates from E. M. Standish, "Keplerian Elements for Approximate Positions of the Major Planets" (JPL, valid
 * 1800-2050); epoch '2019 DECEMBER 12 18:00:00.0' UTC (:396).  Documented deviation, as for the LEO Sun. */
void orc_sun_from_mars(double t, double r[3], double v[3], double *j2000_et)
{
    const double D2R = PI_D / 180.0, AUm = 149597870700.0;
    double days = 7285.25 + 69.184 / 86400.0 + t / 86400.0;       /* TT days from J2000 */
    double T = days / 36525.0;
    double a = (1.52371034 + 0.00001847 * T) * AUm, e = 0.09339410 + 0.00007882 * T;
    double I = (1.84969142 - 0.00813131 * T) * D2R, L = (-4.55343205 + 19140.30268499 * T) * D2R;
    double wbar = (-23.94362959 + 0.44441088 * T) * D2R, Om = (49.55953891 - 0.29257343 * T) * D2R;
    double w = wbar - Om, M = L - wbar;
    double n = (19140.30268499 - 0.44441088) * D2R / (36525.0 * 86400.0);   /* dM/dt [rad/s] */
    M = fmod(M, 2.0 * PI_D);
    double E = M + e * sin(M);
    for (int it = 0; it < 8; it++) E = E - (E - e * sin(E) - M) / (1.0 - e * cos(E));
    double b = a * sqrt(1.0 - e * e);
    double xp = a * (cos(E) - e), yp = b * sin(E);
    double Edot = n / (1.0 - e * cos(E));
    double xd = -a * sin(E) * Edot, yd = b * cos(E) * Edot;
    double cw = cos(w), sw = sin(w), cO = cos(Om), sO = sin(Om), cI = cos(I), sI = sin(I);
    double P1[3] = {cw * cO - sw * sO * cI, cw * sO + sw * cO * cI, sw * sI};
    double P2[3] = {-sw * cO - cw * sO * cI, -sw * sO + cw * cO * cI, cw * sI};
    double eps = 23.43928 * D2R, ce = cos(eps), se = sin(eps);
    double re[3], ve[3];
    for (int k = 0; k < 3; k++) { re[k] = P1[k] * xp + P2[k] * yp; ve[k] = P1[k] * xd + P2[k] * yd; }
    /* ecliptic -> equatorial, heliocentric Mars -> Mars-centred Sun */
    r[0] = -re[0]; r[1] = -(ce * re[1] - se * re[2]); r[2] = -(se * re[1] + ce * re[2]);
    v[0] = -ve[0]; v[1] = -(ce * ve[1] - se * ve[2]); v[2] = -(se * ve[1] + ce * ve[2]);
    if (j2000_et) *j2000_et = days * 86400.0;
}

/* ================================ message bus ================================================= */
typedef struct { uint64_t write_ns; uint64_t count; } MsgHdr;
static void msg_stamp(MsgHdr *h, uint64_t now) { h->write_ns = now; h->count++; }
static int msg_written(const MsgHdr *h) { return h->count > 0; } /* ReadMessage returns false if never written */

typedef struct { MsgHdr h; double r_BN_N[3], v_BN_N[3], sigma_BN[3], omega_BN_B[3]; } SCPlusStatesMsg;
typedef struct { MsgHdr h; double J2000Current, PositionVector[3], VelocityVector[3];
                 double J20002Pfix[3][3], J20002Pfix_dot[3][3]; int computeOrient; } SpicePlanetStateMsg;
typedef struct { MsgHdr h; double neutralDensity; } AtmoPropsMsg;
typedef struct { MsgHdr h; double sigma_BN[3], omega_BN_B[3], vehSunPntBdy[3]; } NavAttMsg;
typedef struct { MsgHdr h; double r_BN_N[3], v_BN_N[3]; } NavTransMsg;
typedef struct { MsgHdr h; double wheelSpeeds[MAX_RW]; } RWSpeedMsg;
typedef struct { MsgHdr h; double motorTorque[MAX_RW]; } RWArrayTorqueMsg;
typedef struct { MsgHdr h; double shadowFactor; } EclipseMsg;
typedef struct { MsgHdr h; double netPower; } PowerNodeUsageMsg;
typedef struct { MsgHdr h; double storageCapacity, storageLevel, currentNetPower; } PowerStorageStatusMsg;
typedef struct { MsgHdr h; double sigma_RN[3], omega_RN_N[3], domega_RN_N[3]; } AttRefMsg;
typedef struct { MsgHdr h; double sigma_BR[3], omega_BR_B[3], omega_RN_B[3], domega_RN_B[3]; } AttGuidMsg;
typedef struct { MsgHdr h; double torqueRequestBody[3]; } CmdTorqueBodyMsg;
typedef struct { MsgHdr h; double thrForce[MAX_THR]; } THRArrayCmdForceMsg;
typedef struct { MsgHdr h; double OnTimeRequest[MAX_THR]; } THRArrayOnTimeCmdMsg;

/* ================================ module state ================================================ */
typedef struct { /* [BSK: thrusterDynamicEffector] one THRConfigSimMsg + THROperationSimMsg */
    double thrLoc_B[3], thrDir_B[3], MaxThrust, MinOnTime;
    double ThrustOnCmd, ThrusterStartTime, ThrustFactor, PreviousIterTime;
    int64_t fireCounter;
} Thruster;

typedef struct { double gsHat_B[3], Js, u_max, u_min, Omega_max, Omega, u_current; } RWheel;

struct orc_leo_sim;
typedef void (*ModelFn)(struct orc_leo_sim *, uint64_t);

typedef struct { /* sys_model_task */
    uint64_t period, next_start;
    int active, n_models;
    ModelFn models[8];
} Task;
typedef struct { Task *task; uint64_t next; int prio; } SchedEntry;

#define MAX_LOG_BYTES 128
typedef struct { /* messageLogger entry */
    const MsgHdr *hdr; const void *payload; size_t bytes;
    uint64_t last_write_check, last_log_time, write_delta;
    unsigned char last[MAX_LOG_BYTES];
    int have;
} LogEntry;

enum { T_DYN, T_SPICE, T_ENV, T_SUNPOINT, T_NADIRPOINT, T_MRPCONTROL, T_RWDESAT,
       T_OPNAVPOINT, T_SUNSAFE, T_MRPRW /* opNav scenario only */, N_TASKS };

struct orc_leo_sim {
    orc_leo_cfg cfg;
    orc_leo_ic ic;
    /* ---- scheduler ---- */
    Task tasks[N_TASKS];
    SchedEntry sched[N_TASKS];
    int n_sched;
    uint64_t next_task_time, current_nanos;
    double simTime;                    /* SIM:78 */
    /* ---- messages ---- */
    SCPlusStatesMsg scState;           /* scObject.scStateOutMsgName */
    SpicePlanetStateMsg sunMsg, earthMsg; /* "sun_planet_data", "earth_planet_data" */
    AtmoPropsMsg atmoMsg;
    NavAttMsg navAtt;
    NavTransMsg navTrans;
    RWSpeedMsg rwSpeeds;
    RWArrayTorqueMsg rwTorqueCmd;      /* "rwTorqueCommand" SIM:307 */
    EclipseMsg eclipseMsg;             /* "eclipse_data_0" SIM:329 */
    PowerNodeUsageMsg panelPower, sinkPower;
    PowerStorageStatusMsg battery;     /* "powerMonitorMsg" */
    AttRefMsg attRef;                  /* "att_reference" */
    AttGuidMsg attGuid;                /* "att_guidance" */
    CmdTorqueBodyMsg cmdTorque;        /* "commandedControlTorque" */
    CmdTorqueBodyMsg deltaH;           /* "wheelDeltaH" */
    THRArrayCmdForceMsg deltaP;        /* "delta_p_achievable" */
    THRArrayOnTimeCmdMsg thrOnTime;    /* "rwDesatTimeOnCmd" */
    /* ---- spacecraftPlus + hub + gravity ---- */
    double mHub, IHub[3][3];
    double r[3], v[3], sigma[3], omega[3];       /* hub states */
    double rDot[3], vDot[3], sigmaDot[3], omegaDot[3], OmegaDot[MAX_RW];
    double timePrevious; uint64_t simTimePrevious;
    uint64_t sysTimeNanos;                        /* "systemTime" property */
    int64_t MRPSwitchCount;
    SpicePlanetStateMsg gravSun, gravEarth;       /* GravBodyData::localPlanet (+ header) */
    double g_N[3];
    /* ---- effectors ---- */
    int nRW; RWheel rw[MAX_RW];
    int nThr; Thruster thr[MAX_THR];
    uint64_t thrPrevCommandTime; double thrPrevFireTime;
    double thrNewCmds[MAX_THR];
    int nFacet; double facetArea[8], facetCd[8], facetN[8][3], facetLoc[8][3];
    double dragDensity;                           /* FacetDrag atmoInData.neutralDensity */
    double extTorquePntB_B[3];
    /* ---- atmosphere ---- */
    double planetRadius, baseDensity, scaleHeight;
    /* ---- power ---- */
    double nHat_B[3], panelArea, panelEfficiency, nodePowerOut;
    double storageCapacity, storedCharge, batPreviousTime;
    /* ---- FSW ---- */
    double sigma_R0N[3];
    double ISC_fsw[3][3];
    double K, P, Ki; uint64_t mrpPriorTime;
    double controlAxes_B[9];
    double hs_min; int initRequest;
    int thrForceSign; double tfm_epsilon, tfm_angErrThresh; int tfm_use2ndLoop; double outTorqAngErr;
    int maxCounterValue, thrDumpingCounter; double thrMinFireTime;
    double thrOnTimeRemaining[MAX_THR]; uint64_t dumpPriorTime, lastDeltaHInMsgTime;
    /* ---- opNav scenario (dynamics half of simulators/opNavSimulator.py + opNav_models/) ---- */
    int scenario;                      /* 0 LEO power/attitude, 1 opNav dynamics */
    double mu_central;
    double sigma_R0R[3];               /* attTrackingError.sigma_R0R (camera frame offset in the opNav scenario) */
    double sHatBdyCmd[3];              /* sunSafePoint */
    NavAttMsg sunPointData;            /* "sun_point_data": cssWlsEst output (truth sun heading substitute) */
    int modeCounter, numModes;
    /* ---- logging ---- */
    LogEntry logs[8]; int n_logs;
    double obs[5];
};

/* ================================ scheduler ([BSK: sys_process.cpp, sim_model.cpp]) ============= */
static void sched_insert(orc_leo_sim *s, SchedEntry e)
{ /* SysProcess::scheduleTask: before the first entry that starts later, or at the same time with lower priority */
    int pos = s->n_sched;
    for (int i = 0; i < s->n_sched; i++) {
        if (s->sched[i].next > e.next || (s->sched[i].next == e.next && e.prio > s->sched[i].prio)) { pos = i; break; }
    }
    for (int i = s->n_sched; i > pos; i--) s->sched[i] = s->sched[i - 1];
    s->sched[pos] = e;
    s->n_sched++;
}
static void sched_add_task(orc_leo_sim *s, int id, double period_s, int prio)
{
    Task *t = &s->tasks[id];
    t->period = sec2nano(period_s); t->next_start = 0; t->active = 1; t->n_models = 0;
    SchedEntry e = {t, 0, prio};
    sched_insert(s, e);
}
static void task_add_model(orc_leo_sim *s, int id, ModelFn f) { Task *t = &s->tasks[id]; t->models[t->n_models++] = f; }
static void task_execute(orc_leo_sim *s, Task *t, uint64_t now)
{ /* SysModelTask::ExecuteTaskList: models run only while active; NextStartTime advances regardless */
    for (int i = 0; i < t->n_models && t->active; i++) t->models[i](s, now);
    t->next_start += t->period;
}
static void proc_single_step_next_task(orc_leo_sim *s, uint64_t now)
{ /* SysProcess::singleStepNextTask */
    SchedEntry e = s->sched[0];
    if (e.next > now) { s->next_task_time = e.next; return; }
    task_execute(s, e.task, now);
    for (int i = 1; i < s->n_sched; i++) s->sched[i - 1] = s->sched[i];
    s->n_sched--;
    e.next = e.task->next_start;
    sched_insert(s, e);
    s->next_task_time = s->sched[0].next;
}
static void log_all_messages(orc_leo_sim *s);
static void sim_single_step_processes(orc_leo_sim *s)
{ /* SimModel::SingleStepProcesses: run every task due at NextTaskTime, then log */
    s->current_nanos = s->next_task_time;
    while (s->next_task_time <= s->current_nanos) proc_single_step_next_task(s, s->current_nanos);
    log_all_messages(s);
}
static void sim_step_until_stop(orc_leo_sim *s, uint64_t stop)
{ /* SimModel::StepUntilStop with stopPri=-1: tasks AT the stop time run (inclusive) */
    while (s->next_task_time <= stop) sim_single_step_processes(s);
}

/* ================================ message logger ([BSK: message_logger.cpp]) ==================== */
static void log_add(orc_leo_sim *s, const MsgHdr *hdr, const void *payload, size_t bytes, uint64_t period)
{
    LogEntry *l = &s->logs[s->n_logs++];
    memset(l, 0, sizeof(*l));
    l->hdr = hdr; l->payload = payload; l->bytes = bytes; l->write_delta = period;
    l->last_log_time = 0xFFFFFFFFFFFFFFFFull;
}
static void log_all_messages(orc_leo_sim *s)
{
    for (int i = 0; i < s->n_logs; i++) {
        LogEntry *l = &s->logs[i];
        int bufferNew = l->last_write_check != l->hdr->count;
        if (bufferNew)
            bufferNew = (l->last_log_time == 0xFFFFFFFFFFFFFFFFull) || (l->hdr->write_ns >= l->last_log_time + l->write_delta);
        l->last_write_check = l->hdr->count;
        if (bufferNew) { memcpy(l->last, l->payload, l->bytes); l->last_log_time = l->hdr->write_ns; l->have = 1; }
    }
}

/* ================================ SURVEY 8(f)-4: ephemeris tables, planet-fixed degree-2 field ======
 * (a) Chebyshev ephemeris tables in the layout of SPICE SPK type 2 / binary PCK type 2 records: n_seg
 *     segments of equal length, three components with n_coef Chebyshev coefficients each; the value is
 *     sum a_k T_k(s), s = (t - mid) / (len/2), the rate is the derivative of the same polynomial (what
 *     spkezr / sxform return for these record types).  Test infrastructure: the tables are process-global.
 *     kind 0: Sun position relative to Earth [m]; kind 1: Earth orientation angles (RA, DEC, W) [rad].
 * (b) Earth orientation without a table: IAU rotation model as in the SPICE text kernel pck00010.tpc
 *     (BODY399_POLE_RA = 0 - 0.641 T, POLE_DEC = 90 - 0.557 T, PM = 190.147 + 360.9856235 d), J2000 -> IAU_EARTH
 *     = R3(W) R1(pi/2 - DEC) R3(pi/2 + RA) (pxform_c), rate from the product rule (sxform_c).
 * (c) [BSK: gravityEffector.cpp GravBodyData::computeGravityInertial + sphericalHarmonics::computeField]:
 *     dcm_PfixN = J20002Pfix + J20002Pfix_dot * dt (dt since the SPICE message was written), Pines' recursion with
 *     normalised coefficients in the planet-fixed frame, rotated back with the transpose.  Restated here for
 *     max degree 2; confidence M (recalled).  Pinned by tests/test_oracle_physics.py against the closed-form
 *     gradient of the degree-2 potential. */
typedef struct { int nseg, ncoef; double t0, seg_len; double *coef; } EphTable;
static EphTable g_eph[3];   /* kind 2: Sun relative to the Mars barycentre [m] (opNav scenario) */
static double g_cbar[5] = {-4.8416537173459064e-04, -2.0661550900e-10, 1.3844138138e-09, 2.4393836573e-06, -1.4002737040e-06};
                           /* C20 = -J2_EARTH / sqrt 5 (the J2 of use_j2); C21, S21, C22, S22: GGM03S-class values, confidence L;
                              orc_set_gravity_coeffs replaces them */
int orc_set_ephemeris(int kind, double t0, double seg_len, int nseg, int ncoef, const double *coef)
{
    if (kind < 0 || kind > 2) return -1;
    free(g_eph[kind].coef); memset(&g_eph[kind], 0, sizeof(EphTable));
    if (nseg <= 0) return 0;
    if (ncoef < 1 || !(seg_len > 0) || !coef) return -1;
    g_eph[kind].coef = (double *)malloc(sizeof(double) * (size_t)nseg * 3 * (size_t)ncoef);
    memcpy(g_eph[kind].coef, coef, sizeof(double) * (size_t)nseg * 3 * (size_t)ncoef);
    g_eph[kind].nseg = nseg; g_eph[kind].ncoef = ncoef; g_eph[kind].t0 = t0; g_eph[kind].seg_len = seg_len;
    return 0;
}
void orc_set_gravity_coeffs(const double cbar[5]) { for (int k = 0; k < 5; k++) g_cbar[k] = cbar[k]; }
int orc_eph_eval(int kind, double t, double val[3], double rate[3])
{ /* forward recurrences T_{k+1} = 2 s T_k - T_{k-1}, U_{k+1} = 2 s U_k - U_{k-1}, T_k' = k U_{k-1} */
    const EphTable *E = &g_eph[kind];
    if (E->nseg <= 0) return -1;
    int i = (int)floor((t - E->t0) / E->seg_len);
    if (i < 0) i = 0;
    if (i >= E->nseg) i = E->nseg - 1;
    double half = 0.5 * E->seg_len, mid = E->t0 + (i + 0.5) * E->seg_len, sc = (t - mid) / half;
    for (int c = 0; c < 3; c++) {
        const double *a = E->coef + ((size_t)i * 3 + c) * E->ncoef;
        double Tm = 1.0, T = sc, Um = 0.0, U = 1.0;     /* T_0, T_1, U_{-1}, U_0 */
        double v = a[0], d = 0.0;
        for (int k = 1; k < E->ncoef; k++) {
            v += a[k] * T; d += a[k] * k * U;
            double Tn = 2.0 * sc * T - Tm, Un = 2.0 * sc * U - Um;
            Tm = T; T = Tn; Um = U; U = Un;
        }
        val[c] = v; rate[c] = d / half;
    }
    return 0;
}
static void rot1(double a, double R[3][3]) { double c = cos(a), s = sin(a); double M[3][3] = {{1, 0, 0}, {0, c, s}, {0, -s, c}}; memcpy(R, M, sizeof(M)); }
static void rot3(double a, double R[3][3]) { double c = cos(a), s = sin(a); double M[3][3] = {{c, s, 0}, {-s, c, 0}, {0, 0, 1}}; memcpy(R, M, sizeof(M)); }
static void drot1(double a, double R[3][3]) { double c = cos(a), s = sin(a); double M[3][3] = {{0, 0, 0}, {0, -s, c}, {0, -c, -s}}; memcpy(R, M, sizeof(M)); }
static void drot3(double a, double R[3][3]) { double c = cos(a), s = sin(a); double M[3][3] = {{-s, c, 0}, {-c, -s, 0}, {0, 0, 0}}; memcpy(R, M, sizeof(M)); }
static void mm3(double A[3][3], double B[3][3], double C[3][3])
{
    double T[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T[i][j] = 0; for (int k = 0; k < 3; k++) T[i][j] += A[i][k] * B[k][j]; }
    memcpy(C, T, sizeof(T));
}
void orc_earth_orientation(double t, double P[3][3], double Pdot[3][3])
{
    double ang[3], rate[3];
    if (orc_eph_eval(1, t, ang, rate) != 0) {
        const double D2R = PI_D / 180.0;
        double d = EPOCH_DAYS_TT_FROM_J2000 + t / 86400.0, T = d / 36525.0;
        ang[0] = (0.0 - 0.641 * T) * D2R; ang[1] = (90.0 - 0.557 * T) * D2R; ang[2] = (190.147 + 360.9856235 * d) * D2R;
        rate[0] = -0.641 / 36525.0 / 86400.0 * D2R; rate[1] = -0.557 / 36525.0 / 86400.0 * D2R; rate[2] = 360.9856235 / 86400.0 * D2R;
    }
    double W[3][3], A[3][3], F[3][3], dW[3][3], dA[3][3], dF[3][3], T1[3][3], T2[3][3], T3[3][3];
    double th = 0.5 * PI_D - ang[1], ph = 0.5 * PI_D + ang[0];
    rot3(ang[2], W); rot1(th, A); rot3(ph, F);
    drot3(ang[2], dW); drot1(th, dA); drot3(ph, dF);
    mm3(A, F, T1); mm3(W, T1, P);
    mm3(A, F, T1); mm3(dW, T1, T1);              /* W' A F * Wdot */
    mm3(dA, F, T2); mm3(W, T2, T2);              /* W A' F * thdot, thdot = -DECdot */
    mm3(A, dF, T3); mm3(W, T3, T3);              /* W A F' * phdot, phdot = RAdot */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        Pdot[i][j] = rate[2] * T1[i][j] - rate[1] * T2[i][j] + rate[0] * T3[i][j];
}
static double sh_getK(int m) { return m == 0 ? 1.0 : 2.0; }
void orc_grav_degree2_pfix(double mu, double radEquator, const double cbar[5], const double pos[3], double acc[3])
{ /* sphericalHarmonics::computeField(pos_Pfix, degree = 2, include_zero_degree = false), Pines' formulation */
    enum { N = 2 };
    double cB[N + 2][N + 2] = {{0}}, sB[N + 2][N + 2] = {{0}}, aBar[N + 3][N + 3] = {{0}};
    double n1[N + 3][N + 3] = {{0}}, n2[N + 3][N + 3] = {{0}}, nq1[N + 3][N + 3] = {{0}}, nq2[N + 3][N + 3] = {{0}};
    cB[0][0] = 1.0; cB[2][0] = cbar[0]; cB[2][1] = cbar[1]; sB[2][1] = cbar[2]; cB[2][2] = cbar[3]; sB[2][2] = cbar[4];
    /* initializeParameters */
    for (int l = 0; l <= N + 1; l++) {
        if (l == 0) aBar[l][l] = 1.0;
        else aBar[l][l] = sqrt((2.0 * l + 1.0) * sh_getK(l) / (2.0 * l * sh_getK(l - 1))) * aBar[l - 1][l - 1];
        for (int m = 0; m <= l; m++) {
            if (l >= m + 2) {
                n1[l][m] = sqrt((2.0 * l + 1.0) * (2.0 * l - 1.0) / ((double)(l - m) * (l + m)));
                n2[l][m] = sqrt((double)(l + m - 1) * (2.0 * l + 1.0) * (l - m - 1) / ((double)(l + m) * (l - m) * (2.0 * l - 3.0)));
            }
        }
    }
    for (int l = 0; l <= N; l++)
        for (int m = 0; m <= l; m++) {
            if (m < l) nq1[l][m] = sqrt((double)(l - m) * sh_getK(m) * (l + m + 1) / sh_getK(m + 1));
            nq2[l][m] = sqrt((double)(l + m + 2) * (l + m + 1) * (2.0 * l + 1.0) * sh_getK(m) / ((2.0 * l + 3.0) * sh_getK(m + 1)));
        }
    /* computeField */
    double x = pos[0], y = pos[1], z = pos[2];
    double r = sqrt(x * x + y * y + z * z), sx = x / r, ty = y / r, u = z / r;
    for (int l = 1; l <= N + 1; l++) aBar[l][l - 1] = sqrt((2.0 * l) * sh_getK(l - 1) / sh_getK(l)) * aBar[l][l] * u;
    for (int m = 0; m <= N + 1; m++)
        for (int l = m + 2; l <= N + 1; l++) aBar[l][m] = u * n1[l][m] * aBar[l - 1][m] - n2[l][m] * aBar[l - 2][m];
    double rE[N + 2], iM[N + 2], rhol[N + 3];
    rE[0] = 1.0; iM[0] = 0.0;
    for (int m = 1; m <= N + 1; m++) { rE[m] = sx * rE[m - 1] - ty * iM[m - 1]; iM[m] = sx * iM[m - 1] + ty * rE[m - 1]; }
    double rho = radEquator / r;
    rhol[0] = mu / r; rhol[1] = rhol[0] * rho;
    double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    for (int l = 1; l <= N; l++) {       /* degree 1 carries zero coefficients */
        rhol[l + 1] = rho * rhol[l];
        double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (int m = 0; m <= l; m++) {
            double D = cB[l][m] * rE[m] + sB[l][m] * iM[m], E = 0, F = 0;
            if (m > 0) { E = cB[l][m] * rE[m - 1] + sB[l][m] * iM[m - 1]; F = sB[l][m] * rE[m - 1] - cB[l][m] * iM[m - 1]; }
            s1 += m * aBar[l][m] * E; s2 += m * aBar[l][m] * F;
            if (m < l) s3 += nq1[l][m] * aBar[l][m + 1] * D;
            s4 += nq2[l][m] * aBar[l + 1][m + 1] * D;
        }
        a1 += rhol[l + 1] / radEquator * s1; a2 += rhol[l + 1] / radEquator * s2;
        a3 += rhol[l + 1] / radEquator * s3; a4 -= rhol[l + 1] / radEquator * s4;
    }
    acc[0] = a1 + sx * a4; acc[1] = a2 + ty * a4; acc[2] = a3 + u * a4;
}
static void grav_degree2_pfix(const SpicePlanetStateMsg *b, uint64_t systemClock, const double r_I[3], double out[3])
{
    double dt = (double)(systemClock - b->h.write_ns) * NANO2SEC;      /* unsigned clock difference, as the body position */
    double D[3][3], rp[3], gp[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D[i][j] = b->J20002Pfix[i][j] + b->J20002Pfix_dot[i][j] * dt;
    for (int i = 0; i < 3; i++) rp[i] = D[i][0] * r_I[0] + D[i][1] * r_I[1] + D[i][2] * r_I[2];
    orc_grav_degree2_pfix(MU_EARTH, REQ_EARTH_KM * 1000.0, g_cbar, rp, gp);
    for (int j = 0; j < 3; j++) out[j] = D[0][j] * gp[0] + D[1][j] * gp[1] + D[2][j] * gp[2];
}

/* ================================ SPICE stand-in ([BSK: spice_interface.cpp]) =================== */
static void spice_update(orc_leo_sim *s, uint64_t now)
{ /* zeroBase = "earth" (SIM:225): Earth at the origin, Sun relative to Earth */
    double et;
    if (s->scenario == 1) orc_sun_from_mars(now * NANO2SEC, s->sunMsg.PositionVector, s->sunMsg.VelocityVector, &et);
    else {
        orc_sun_ephemeris(now * NANO2SEC, s->sunMsg.PositionVector, s->sunMsg.VelocityVector, &et);
        orc_eph_eval(0, now * NANO2SEC, s->sunMsg.PositionVector, s->sunMsg.VelocityVector);   /* table, when one is loaded */
        if (s->cfg.grav_pfix) {   /* computeOrient: pxform_c / sxform_c at the message time */
            orc_earth_orientation(now * NANO2SEC, s->earthMsg.J20002Pfix, s->earthMsg.J20002Pfix_dot);
            s->earthMsg.computeOrient = 1;
        }
    }
    s->sunMsg.J2000Current = et;
    s->earthMsg.J2000Current = et;
    v3SetZero(s->earthMsg.PositionVector); v3SetZero(s->earthMsg.VelocityVector);
    msg_stamp(&s->sunMsg.h, now); msg_stamp(&s->earthMsg.h, now);
}

/* ================================ spacecraftPlus ================================================ */
static void MRP_toRotationMatrix(const double q[3], double NB[3][3])
{ /* [BSK: avsEigenMRP.h MRPBase::toRotationMatrix] -> dcm_NB */
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    double ps2 = 1 + n2, ms2 = 1 - n2, ms2Sq = ms2 * ms2;
    double s1s2 = 8 * q[0] * q[1], s1s3 = 8 * q[0] * q[2], s2s3 = 8 * q[1] * q[2];
    double s1Sq = q[0] * q[0], s2Sq = q[1] * q[1], s3Sq = q[2] * q[2];
    NB[0][0] = 4 * (s1Sq - s2Sq - s3Sq) + ms2Sq;
    NB[0][1] = s1s2 - 4 * q[2] * ms2;
    NB[0][2] = s1s3 + 4 * q[1] * ms2;
    NB[1][0] = s1s2 + 4 * q[2] * ms2;
    NB[1][1] = 4 * (-s1Sq + s2Sq - s3Sq) + ms2Sq;
    NB[1][2] = s2s3 - 4 * q[0] * ms2;
    NB[2][0] = s1s3 - 4 * q[1] * ms2;
    NB[2][1] = s2s3 + 4 * q[0] * ms2;
    NB[2][2] = 4 * (-s1Sq - s2Sq + s3Sq) + ms2Sq;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) NB[i][j] = NB[i][j] / ps2 / ps2;
}
static void grav_body_position(const SpicePlanetStateMsg *b, uint64_t systemClock, double out[3])
{ /* GravityEffector::getEulerSteppedGravBodyPosition */
    double dt = (double)(systemClock - b->h.write_ns) * NANO2SEC;
    for (int k = 0; k < 3; k++) out[k] = b->PositionVector[k] + b->VelocityVector[k] * dt;
}
static void grav_point_mass(double mu, const double r_I[3], double out[3])
{ /* GravBodyData::computeGravityInertial: -r*mu/|r|^3 */
    double rMag = v3Norm(r_I);
    for (int k = 0; k < 3; k++) out[k] = -r_I[k] * mu / (rMag * rMag * rMag);
}
static void grav_j2(const double r_I[3], double out[3])
{ /* degree-2 zonal only, planet-fixed frame == inertial (computeOrient off); stress config only */
    double r2 = v3Dot(r_I, r_I), r = sqrt(r2), Re = REQ_EARTH_KM * 1000.0;
    double k = -1.5 * J2_EARTH * MU_EARTH * Re * Re / (r2 * r2 * r);
    double z2 = 5.0 * r_I[2] * r_I[2] / r2;
    out[0] = k * r_I[0] * (1.0 - z2);
    out[1] = k * r_I[1] * (1.0 - z2);
    out[2] = k * r_I[2] * (3.0 - z2);
}
static void gravity_compute(orc_leo_sim *s, const double r_cF_N[3])
{ /* GravityEffector::computeGravityField; Earth is central, Sun is a third body (SIM:227-232) */
    double r_CN_N[3], r_cN_N[3], r_PN_N[3], r_cP_N[3], tmp[3], d[3], acc[3] = {0, 0, 0};
    grav_body_position(&s->gravEarth, s->sysTimeNanos, r_CN_N);
    v3Add(r_cF_N, r_CN_N, r_cN_N);
    /* gravBodies dict order: sun created first (SIM:227), then earth (SIM:228) */
    grav_body_position(&s->gravSun, s->sysTimeNanos, r_PN_N);
    v3Subtract(r_cN_N, r_PN_N, r_cP_N);
    v3Subtract(r_PN_N, r_CN_N, d);
    grav_point_mass(MU_SUN, d, tmp); v3Add(acc, tmp, acc);
    grav_point_mass(MU_SUN, r_cP_N, tmp); v3Add(acc, tmp, acc);
    grav_body_position(&s->gravEarth, s->sysTimeNanos, r_PN_N);
    v3Subtract(r_cN_N, r_PN_N, r_cP_N);
    grav_point_mass(s->mu_central, r_cP_N, tmp);
    if (s->cfg.grav_pfix) { double j[3]; grav_degree2_pfix(&s->gravEarth, s->sysTimeNanos, r_cP_N, j); v3Add(tmp, j, tmp); }
    else if (s->cfg.use_j2) { double j[3]; grav_j2(r_cP_N, j); v3Add(tmp, j, tmp); }
    v3Add(acc, tmp, acc);
    v3Copy(acc, s->g_N);
}
static void drag_compute(orc_leo_sim *s, double F_B[3], double L_B[3])
{ /* FacetDragDynamicEffector::computeForceTorque = updateDragDir + plateDrag */
    double NB[3][3], v_B[3], v_hat_B[3];
    MRP_toRotationMatrix(s->sigma, NB);
    m33tMultV3(NB, s->v, v_B);
    double vn = v3Norm(v_B);
    for (int k = 0; k < 3; k++) v_hat_B[k] = v_B[k] / vn;
    v3SetZero(F_B); v3SetZero(L_B);
    for (int i = 0; i < s->nFacet; i++) {
        double projectionTerm = v3Dot(s->facetN[i], v_hat_B);
        double projectedArea = s->facetArea[i] * projectionTerm;
        if (projectedArea > 0.0) {
            double c = 0.5 * pow(vn, 2.0) * s->facetCd[i] * projectedArea * s->dragDensity * (-1.0);
            double f[3], t[3];
            v3Scale(c, v_hat_B, f);
            v3Cross(f, s->facetLoc[i], t);
            v3Scale(-1., t, t);
            v3Add(F_B, f, F_B);
            v3Add(L_B, t, L_B);
        }
    }
}
static void thrusters_compute(orc_leo_sim *s, double integTime, double F_B[3], double L_B[3])
{ /* ThrusterDynamicEffector::computeForceTorque, no ramps: ThrustFactor is 0 or 1 */
    double dt = integTime - s->thrPrevFireTime;
    v3SetZero(F_B); v3SetZero(L_B);
    for (int i = 0; i < s->nThr; i++) {
        Thruster *t = &s->thr[i];
        if ((t->ThrustOnCmd + t->ThrusterStartTime - integTime) >= -dt * 10E-10 && t->ThrustOnCmd > 0.0) {
            t->PreviousIterTime = integTime;          /* ComputeThrusterFire */
            t->ThrustFactor = 1.0;
        } else if (t->ThrustFactor > 0.0) {
            t->ThrustFactor = 0.0;                    /* ComputeThrusterShut */
        }
        double mag = t->MaxThrust * t->ThrustFactor, f[3], l[3];
        v3Scale(mag, t->thrDir_B, f);
        v3Add(f, F_B, F_B);
        v3Cross(t->thrLoc_B, f, l);
        v3Add(l, L_B, L_B);
    }
    s->thrPrevFireTime = integTime;
}
static void sc_equations_of_motion(orc_leo_sim *s, double t)
{ /* SpacecraftPlus::equationsOfMotion, c_B = 0 (balanced wheels carry no mass properties) */
    s->sysTimeNanos = (uint64_t)((double)s->simTimePrevious + (t - s->timePrevious) / NANO2SEC);
    gravity_compute(s, s->r);
    /* dynamic effectors in attach order: drag (SIM:284), extForceTorque (SIM:298), thrusters (SIM:318) */
    double sumF_B[3] = {0, 0, 0}, sumL_B[3] = {0, 0, 0}, f[3], l[3];
    drag_compute(s, f, l); v3Add(sumF_B, f, sumF_B); v3Add(sumL_B, l, sumL_B);
    v3Add(sumL_B, s->extTorquePntB_B, sumL_B);
    thrusters_compute(s, t, f, l); v3Add(sumF_B, f, sumF_B); v3Add(sumL_B, l, sumL_B);
    /* state effector back-substitution: ReactionWheelStateEffector::updateContributions (BalancedWheels) */
    double D[3][3] = {{0}}, vecRot[3] = {0, 0, 0};
    for (int i = 0; i < s->nRW; i++) {
        RWheel *w = &s->rw[i];
        double wxg[3];
        v3Cross(s->omega, w->gsHat_B, wxg);
        for (int a = 0; a < 3; a++) {
            for (int b = 0; b < 3; b++) D[a][b] -= w->Js * w->gsHat_B[a] * w->gsHat_B[b];
            vecRot[a] -= w->gsHat_B[a] * w->u_current + w->Js * w->Omega * wxg[a];
        }
    }
    /* hub contributions (HubEffector / SpacecraftPlus) */
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) D[a][b] += s->IHub[a][b];
    double Iw[3], wxIw[3];
    m33MultV3(s->IHub, s->omega, Iw);
    v3Cross(s->omega, Iw, wxIw);
    for (int a = 0; a < 3; a++) vecRot[a] += -wxIw[a] + sumL_B[a];
    /* HubEffector::computeDerivatives with matrixB = matrixC = 0, matrixA = m I */
    double Dinv[3][3], NB[3][3], a_B[3], a_N[3];
    m33Inverse(D, Dinv);
    m33MultV3(Dinv, vecRot, s->omegaDot);
    MRP_toRotationMatrix(s->sigma, NB);
    for (int k = 0; k < 3; k++) a_B[k] = sumF_B[k] / s->mHub;
    m33MultV3(NB, a_B, a_N);
    v3Add(a_N, s->g_N, s->vDot);
    v3Copy(s->v, s->rDot);
    { /* sigmaDot = 1/4 [B(sigma)] omega */
        const double *q = s->sigma, *w = s->omega;
        double n2 = v3Dot(q, q), B[3][3];
        B[0][0] = 1 - n2 + 2 * q[0] * q[0]; B[0][1] = 2 * (q[0] * q[1] - q[2]); B[0][2] = 2 * (q[0] * q[2] + q[1]);
        B[1][0] = 2 * (q[1] * q[0] + q[2]); B[1][1] = 1 - n2 + 2 * q[1] * q[1]; B[1][2] = 2 * (q[1] * q[2] - q[0]);
        B[2][0] = 2 * (q[2] * q[0] - q[1]); B[2][1] = 2 * (q[2] * q[1] + q[0]); B[2][2] = 1 - n2 + 2 * q[2] * q[2];
        double Bw[3];
        m33MultV3(B, w, Bw);
        v3Scale(1.0 / 4.0, Bw, s->sigmaDot);
    }
    /* ReactionWheelStateEffector::computeDerivatives */
    for (int i = 0; i < s->nRW; i++)
        s->OmegaDot[i] = s->rw[i].u_current / s->rw[i].Js - v3Dot(s->rw[i].gsHat_B, s->omegaDot);
}
#define NSTATE (12 + MAX_RW)
static void sc_pack(const orc_leo_sim *s, double x[NSTATE])
{
    memcpy(x, s->r, 24); memcpy(x + 3, s->v, 24); memcpy(x + 6, s->sigma, 24); memcpy(x + 9, s->omega, 24);
    for (int i = 0; i < MAX_RW; i++) x[12 + i] = s->rw[i].Omega;
}
static void sc_unpack(orc_leo_sim *s, const double x[NSTATE])
{
    memcpy(s->r, x, 24); memcpy(s->v, x + 3, 24); memcpy(s->sigma, x + 6, 24); memcpy(s->omega, x + 9, 24);
    for (int i = 0; i < MAX_RW; i++) s->rw[i].Omega = x[12 + i]; /* updateEffectorMassProps refreshes RW.Omega */
}
static void sc_deriv(const orc_leo_sim *s, double k[NSTATE])
{
    memcpy(k, s->rDot, 24); memcpy(k + 3, s->vDot, 24); memcpy(k + 6, s->sigmaDot, 24); memcpy(k + 9, s->omegaDot, 24);
    for (int i = 0; i < MAX_RW; i++) k[12 + i] = i < s->nRW ? s->OmegaDot[i] : 0.0;
}
static void sc_integrate_rk4(orc_leo_sim *s, double currentTime, double h)
{ /* [BSK: svIntegratorRK4::integrate] */
    double xInit[NSTATE], xOut[NSTATE], x[NSTATE], k[NSTATE];
    sc_pack(s, xInit); memcpy(xOut, xInit, sizeof(xOut));
    sc_equations_of_motion(s, currentTime);
    sc_deriv(s, k);
    for (int i = 0; i < NSTATE; i++) { xOut[i] += k[i] * (h / 6.0); x[i] = xInit[i] + 0.5 * h * k[i]; }
    sc_unpack(s, x);
    sc_equations_of_motion(s, currentTime + h * 0.5);
    sc_deriv(s, k);
    for (int i = 0; i < NSTATE; i++) { xOut[i] += k[i] * (h / 3.0); x[i] = xInit[i] + 0.5 * h * k[i]; }
    sc_unpack(s, x);
    sc_equations_of_motion(s, currentTime + h * 0.5);
    sc_deriv(s, k);
    for (int i = 0; i < NSTATE; i++) { xOut[i] += k[i] * (h / 3.0); x[i] = xInit[i] + h * k[i]; }
    sc_unpack(s, x);
    sc_equations_of_motion(s, currentTime + h);
    sc_deriv(s, k);
    for (int i = 0; i < NSTATE; i++) xOut[i] += k[i] * (h / 6.0);
    sc_unpack(s, xOut);
}
static void sc_update(orc_leo_sim *s, uint64_t now)
{ /* SpacecraftPlus::UpdateState */
    double newTime = now * NANO2SEC;
    /* gravField.UpdateState -> GravBodyData::loadEphemeris (copy of the SPICE messages + headers) */
    if (msg_written(&s->sunMsg.h)) s->gravSun = s->sunMsg;
    if (msg_written(&s->earthMsg.h)) s->gravEarth = s->earthMsg;
    /* integrateState */
    double localTimeStep = newTime - s->timePrevious;
    double timeBefore = newTime - localTimeStep;
    sc_integrate_rk4(s, timeBefore, localTimeStep);
    s->timePrevious = newTime;
    /* HubEffector::modifyStates: MRP shadow-set switch */
    if (v3Norm(s->sigma) > 1) {
        double d = v3Dot(s->sigma, s->sigma);
        for (int k = 0; k < 3; k++) s->sigma[k] = -s->sigma[k] / d;
        s->MRPSwitchCount++;
    }
    /* writeOutputMessages */
    v3Copy(s->r, s->scState.r_BN_N); v3Copy(s->v, s->scState.v_BN_N);
    v3Copy(s->sigma, s->scState.sigma_BN); v3Copy(s->omega, s->scState.omega_BN_B);
    msg_stamp(&s->scState.h, now);
    s->simTimePrevious = now;
}

/* ================================ DynTask / EnvTask modules ==================================== */
static void atmosphere_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: atmosphereBase.cpp + exponentialAtmosphere.cpp]; planet message unset -> planet at origin */
    if (msg_written(&s->scState.h)) {
        double orbitRadius = v3Norm(s->scState.r_BN_N);
        double orbitAltitude = orbitRadius - s->planetRadius;
        s->atmoMsg.neutralDensity = s->baseDensity * exp(-(orbitAltitude) / s->scaleHeight);
    }
    msg_stamp(&s->atmoMsg.h, now);
}
static void drag_update(orc_leo_sim *s, uint64_t now)
{ /* FacetDragDynamicEffector::UpdateState -> ReadInputs (latch density) */
    (void)now;
    if (msg_written(&s->atmoMsg.h)) s->dragDensity = s->atmoMsg.neutralDensity;
}
static void simple_nav_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: simple_nav.cpp] with the default zero PMatrix/walkBounds -> estimate == truth (SIM:321-323) */
    double sc2Sun[3], BN[3][3];
    v3Copy(s->scState.r_BN_N, s->navTrans.r_BN_N); v3Copy(s->scState.v_BN_N, s->navTrans.v_BN_N);
    v3Copy(s->scState.sigma_BN, s->navAtt.sigma_BN); v3Copy(s->scState.omega_BN_B, s->navAtt.omega_BN_B);
    v3Subtract(s->sunMsg.PositionVector, s->scState.r_BN_N, sc2Sun);
    v3Normalize(sc2Sun, sc2Sun);
    orc_MRP2C(s->scState.sigma_BN, BN);
    m33MultV3(BN, sc2Sun, s->navAtt.vehSunPntBdy);
    msg_stamp(&s->navAtt.h, now); msg_stamp(&s->navTrans.h, now);
}
static void rw_update(orc_leo_sim *s, uint64_t now)
{ /* ReactionWheelStateEffector::UpdateState = ReadInputs + ConfigureRWRequests + WriteOutputMessages */
    for (int i = 0; i < s->nRW; i++) {
        RWheel *w = &s->rw[i];
        double u_cmd = msg_written(&s->rwTorqueCmd.h) ? s->rwTorqueCmd.motorTorque[i] : 0.0;
        if (w->u_max > 0) { if (u_cmd > w->u_max) u_cmd = w->u_max; else if (u_cmd < -w->u_max) u_cmd = -w->u_max; }
        if (fabs(u_cmd) < w->u_min) u_cmd = 0.0;
        if (fabs(w->Omega) >= w->Omega_max && w->Omega_max > 0.0 && w->Omega * u_cmd >= 0.0) u_cmd = 0.0;
        w->u_current = u_cmd;
        s->rwSpeeds.wheelSpeeds[i] = w->Omega;
    }
    msg_stamp(&s->rwSpeeds.h, now);
}
static void thruster_update(orc_leo_sim *s, uint64_t now)
{ /* ThrusterDynamicEffector::UpdateState: act only on a NEW on-time message */
    (void)now;
    if (!msg_written(&s->thrOnTime.h) || s->thrPrevCommandTime == s->thrOnTime.h.write_ns) return;
    s->thrPrevCommandTime = s->thrOnTime.h.write_ns;
    double currentTime = s->thrPrevCommandTime * 1.0E-9;
    for (int i = 0; i < s->nThr; i++) { /* ConfigureThrustRequests */
        Thruster *t = &s->thr[i];
        double cmd = s->thrOnTime.OnTimeRequest[i];
        if (cmd >= t->MinOnTime) {
            t->ThrustOnCmd = cmd;
            t->fireCounter += t->ThrustFactor > 0.0 ? 0 : 1;
        } else {
            t->ThrustOnCmd = t->ThrustFactor > 0.0 ? cmd : 0.0;
        }
        t->ThrusterStartTime = currentTime;
        t->PreviousIterTime = currentTime;
    }
}
double orc_eclipse_shadow(const double r_HN_N[3], const double r_PN_N[3], const double r_BN_N[3], double planetRadius)
{ /* [BSK: eclipse.cpp UpdateState + computePercentShadow], one planet */
    double s_HP_N[3], r_HB_N[3], s_BP_N[3];
    double shadow = 1.0;
    v3Subtract(r_HN_N, r_PN_N, s_HP_N);
    v3Subtract(r_HN_N, r_BN_N, r_HB_N);
    v3Subtract(r_BN_N, r_PN_N, s_BP_N);
    if (v3Norm(r_HB_N) < v3Norm(s_HP_N)) return shadow;     /* spacecraft in front of the planet */
    double sn = v3Norm(s_BP_N), hp = v3Norm(s_HP_N);
    double RS = REQ_SUN_KM * 1000;
    double f_1 = asin((RS + planetRadius) / hp);
    double f_2 = asin((RS - planetRadius) / hp);
    double s_0 = (-v3Dot(s_BP_N, s_HP_N)) / hp;
    double c_1 = s_0 + planetRadius / sin(f_1);
    double c_2 = s_0 - planetRadius / sin(f_2);
    double l = sqrt(sn * sn - s_0 * s_0);
    double l_1 = c_1 * tan(f_1);
    double l_2 = c_2 * tan(f_2);
    if (fabs(l) < fabs(l_2) || fabs(l) < fabs(l_1)) {
        /* total / annular / partial all go through computePercentShadow */
        double normR_HB_N = v3Norm(r_HB_N), normS_BP_N = sn;
        double a = safeAsin(RS / normR_HB_N);
        double b = safeAsin(planetRadius / normS_BP_N);
        double c = safeAcos((-v3Dot(s_BP_N, r_HB_N)) / (normS_BP_N * normR_HB_N));
        if (c < b - a) {
            shadow = 0.0;
        } else if (c < a - b) {
            double areaSun = PI_D * a * a, areaBody = PI_D * b * b;
            double area = areaSun - areaBody;
            shadow = 1 - area / (PI_D * a * a);
        } else if (c < a + b) {
            double x = (c * c + a * a - b * b) / (2 * c);
            double y = sqrt(a * a - x * x);
            double area = a * a * safeAcos(x / a) + b * b * safeAcos((c - x) / b) - c * y;
            shadow = 1 - area / (PI_D * a * a);
        }
    }
    return shadow;
}
static void eclipse_update(orc_leo_sim *s, uint64_t now)
{
    s->eclipseMsg.shadowFactor = orc_eclipse_shadow(s->sunMsg.PositionVector, s->earthMsg.PositionVector,
                                                    s->scState.r_BN_N, REQ_EARTH_KM * 1000);
    msg_stamp(&s->eclipseMsg.h, now);
}
static void solar_panel_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: simpleSolarPanel.cpp] incl. the 1.8 sun-distance factor */
    double r_SB_N[3], sHat_N[3], sHat_B[3], BN[3][3];
    v3Subtract(s->sunMsg.PositionVector, s->scState.r_BN_N, r_SB_N);
    double d = v3Norm(r_SB_N);
    for (int k = 0; k < 3; k++) sHat_N[k] = r_SB_N[k] / d;
    double NB[3][3];
    MRP_toRotationMatrix(s->scState.sigma_BN, NB);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) BN[i][j] = NB[j][i];
    m33MultV3(BN, sHat_N, sHat_B);
    double projectedArea = s->panelArea * v3Dot(sHat_B, s->nHat_B);
    if (projectedArea < 0) projectedArea = 0;
    double sunDistanceFactor = pow(AU_KM * 1000., 2.) / pow(d, 2.);
    double shadowFactor = msg_written(&s->eclipseMsg.h) ? s->eclipseMsg.shadowFactor : 1.0;
    s->panelPower.netPower = s->panelEfficiency * sunDistanceFactor * SOLAR_FLUX_EARTH * projectedArea * shadowFactor;
    msg_stamp(&s->panelPower.h, now);
}
static void battery_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: powerStorageBase.cpp + simpleBattery.cpp]; integrates only if EVERY node message was readable (quirk Q2) */
    if (msg_written(&s->panelPower.h) && msg_written(&s->sinkPower.h)) {
        double currentTime = now * NANO2SEC;
        double currentTimestep = currentTime - s->batPreviousTime;
        double currentPowerSum = 0.0;
        currentPowerSum += s->panelPower.netPower;   /* addPowerNodeToModel order SIM:344-345 */
        currentPowerSum += s->sinkPower.netPower;
        s->storedCharge = s->storedCharge + currentPowerSum * currentTimestep;
        if (s->storedCharge > s->storageCapacity) s->storedCharge = s->storageCapacity;
        if (s->storedCharge < 0) s->storedCharge = 0;
        s->battery.storageCapacity = s->storageCapacity;
        s->battery.storageLevel = s->storedCharge;
        s->battery.currentNetPower = currentPowerSum;
        s->batPreviousTime = currentTime;
    } else {
        s->battery.storageCapacity = 0; s->battery.storageLevel = 0; s->battery.currentNetPower = 0;
    }
    msg_stamp(&s->battery.h, now);
}
static void power_sink_update(orc_leo_sim *s, uint64_t now)
{
    s->sinkPower.netPower = s->nodePowerOut;
    msg_stamp(&s->sinkPower.h, now);
}

/* ================================ FSW modules ================================================== */
static void inertial3D_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: inertial3D.c] */
    v3Copy(s->sigma_R0N, s->attRef.sigma_RN);
    v3SetZero(s->attRef.omega_RN_N); v3SetZero(s->attRef.domega_RN_N);
    msg_stamp(&s->attRef.h, now);
}
static void hillPoint_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: hillPoint.c computeHillPointingReference] */
    double r_BN_N[3] = {0, 0, 0}, v_BN_N[3] = {0, 0, 0}, celPos[3] = {0, 0, 0}, celVel[3] = {0, 0, 0};
    if (msg_written(&s->navTrans.h)) { v3Copy(s->navTrans.r_BN_N, r_BN_N); v3Copy(s->navTrans.v_BN_N, v_BN_N); }
    if (s->cfg.hill_cel_pun && msg_written(&s->earthMsg.h)) {
        /* an EphemerisIntMsg-sized read of a SpicePlanetStateSimMsg aliases
         * r_BdyZero_N = {J2000Current, Pos[0], Pos[1]}, v_BdyZero_N = {Pos[2], Vel[0], Vel[1]} */
        celPos[0] = s->earthMsg.J2000Current; celPos[1] = s->earthMsg.PositionVector[0]; celPos[2] = s->earthMsg.PositionVector[1];
        celVel[0] = s->earthMsg.PositionVector[2]; celVel[1] = s->earthMsg.VelocityVector[0]; celVel[2] = s->earthMsg.VelocityVector[1];
    }
    double relPos[3], relVel[3], dcm_RN[3][3], h[3];
    v3Subtract(r_BN_N, celPos, relPos);
    v3Subtract(v_BN_N, celVel, relVel);
    v3Normalize(relPos, dcm_RN[0]);
    v3Cross(relPos, relVel, h);
    v3Normalize(h, dcm_RN[2]);
    v3Cross(dcm_RN[2], dcm_RN[0], dcm_RN[1]);
    orc_C2MRP(dcm_RN, s->attRef.sigma_RN);
    double rm = v3Norm(relPos), hm = v3Norm(h), dfdt, ddfdt2;
    if (rm > 1.) {
        dfdt = hm / (rm * rm);
        ddfdt2 = -2.0 * v3Dot(relVel, dcm_RN[0]) / rm * dfdt;
    } else { dfdt = 0.; ddfdt2 = 0.; }
    double omega_RN_R[3] = {0, 0, dfdt}, domega_RN_R[3] = {0, 0, ddfdt2};
    m33tMultV3(dcm_RN, omega_RN_R, s->attRef.omega_RN_N);
    m33tMultV3(dcm_RN, domega_RN_R, s->attRef.domega_RN_N);
    msg_stamp(&s->attRef.h, now);
}
static void attTrackingError_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: attTrackingError.c computeAttitudeError], sigma_R0R = 0 */
    AttRefMsg ref; NavAttMsg nav;
    memset(&ref, 0, sizeof(ref)); memset(&nav, 0, sizeof(nav));
    if (msg_written(&s->attRef.h)) ref = s->attRef;
    if (msg_written(&s->navAtt.h)) nav = s->navAtt;
    double sigma_RR0[3], sigma_RN[3], dcm_BN[3][3];
    v3Scale(-1.0, s->sigma_R0R, sigma_RR0);         /* zero in the LEO scenario */
    orc_addMRP(ref.sigma_RN, sigma_RR0, sigma_RN);
    orc_subMRP(nav.sigma_BN, sigma_RN, s->attGuid.sigma_BR);
    orc_MRP2C(nav.sigma_BN, dcm_BN);
    m33MultV3(dcm_BN, ref.omega_RN_N, s->attGuid.omega_RN_B);
    v3Subtract(nav.omega_BN_B, s->attGuid.omega_RN_B, s->attGuid.omega_BR_B);
    m33MultV3(dcm_BN, ref.domega_RN_N, s->attGuid.domega_RN_B);
    msg_stamp(&s->attGuid.h, now);
}
static void MRP_Feedback_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: MRP_Feedback.c]; Ki < 0 -> integral off; no RW message wired (SIM:440-449) -> numRW = 0 */
    AttGuidMsg g; memset(&g, 0, sizeof(g));
    if (msg_written(&s->attGuid.h)) g = s->attGuid;
    s->mrpPriorTime = now;
    double omega_BN_B[3], Lr[3], v3_1[3], v3_2[3], v3[3], v3_4[3], v3_6[3], v3_7[3], v3_8[3], v3_9[3], v3_10[3];
    double z[3] = {0, 0, 0}, known[3] = {0, 0, 0};
    v3Add(g.omega_BR_B, g.omega_RN_B, omega_BN_B);
    v3Scale(s->K, g.sigma_BR, Lr);
    v3Scale(s->P, g.omega_BR_B, v3_1);
    v3Add(v3_1, Lr, Lr);
    v3Scale(s->Ki, z, v3_2);
    v3Scale(s->P, v3_2, v3);
    v3Add(v3, Lr, Lr);
    m33MultV3(s->ISC_fsw, omega_BN_B, v3_4);
    v3Add(g.omega_RN_B, v3_2, v3_6);
    v3Cross(v3_6, v3_4, v3_7);
    v3Subtract(Lr, v3_7, Lr);
    v3Cross(omega_BN_B, g.omega_RN_B, v3_8);
    v3Subtract(g.domega_RN_B, v3_8, v3_9);
    m33MultV3(s->ISC_fsw, v3_9, v3_10);
    v3Subtract(Lr, v3_10, Lr);
    v3Add(known, Lr, Lr);
    v3Scale(-1.0, Lr, Lr);
    v3Copy(Lr, s->cmdTorque.torqueRequestBody);
    msg_stamp(&s->cmdTorque.h, now);
}
static void rwMotorTorque_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: rwMotorTorque.c] minimum-norm map onto the available wheels */
    double Lr_B[3] = {0, 0, 0}, Lr_C[3] = {0, 0, 0}, us[MAX_RW] = {0, 0, 0, 0}, CGs[3][MAX_RW];
    int numControlAxes = 0;
    for (int i = 0; i < 3; i++) if (v3Norm(&s->controlAxes_B[3 * numControlAxes]) > 0.0) numControlAxes++;
    if (msg_written(&s->cmdTorque.h)) v3Copy(s->cmdTorque.torqueRequestBody, Lr_B);
    v3Scale(-1.0, Lr_B, Lr_B);
    for (int i = 0; i < numControlAxes; i++) Lr_C[i] = v3Dot(&s->controlAxes_B[3 * i], Lr_B);
    memset(CGs, 0, sizeof(CGs));
    for (int i = 0; i < numControlAxes; i++)
        for (int j = 0; j < s->nRW; j++) CGs[i][j] = v3Dot(s->rw[j].gsHat_B, &s->controlAxes_B[3 * i]);
    if (s->nRW >= numControlAxes) {
        double M[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, v3_temp[3];
        for (int i = 0; i < numControlAxes; i++)
            for (int j = 0; j < numControlAxes; j++) {
                M[i][j] = 0.0;
                for (int k = 0; k < s->nRW; k++) M[i][j] += CGs[i][k] * CGs[j][k];
            }
        m33Inverse(M, M);
        m33MultV3(M, Lr_C, v3_temp);
        for (int i = 0; i < s->nRW; i++)
            for (int j = 0; j < numControlAxes; j++) us[i] += CGs[j][i] * v3_temp[j];
    }
    for (int i = 0; i < MAX_RW; i++) s->rwTorqueCmd.motorTorque[i] = us[i];
    msg_stamp(&s->rwTorqueCmd.h, now);
}
static void thrMomentumManagement_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: thrMomentumManagement.c]: one-shot after Reset */
    if (s->initRequest == 1) {
        double hs_B[3] = {0, 0, 0}, vec3[3], Delta_H_B[3];
        for (int i = 0; i < s->nRW; i++) {
            double ws = msg_written(&s->rwSpeeds.h) ? s->rwSpeeds.wheelSpeeds[i] : 0.0;
            v3Scale(s->rw[i].Js * ws, s->rw[i].gsHat_B, vec3);
            v3Add(hs_B, vec3, hs_B);
        }
        double hs = v3Norm(hs_B);
        if (hs < s->hs_min) v3SetZero(Delta_H_B);
        else v3Scale(-(hs - s->hs_min) / hs, hs_B, Delta_H_B);
        s->initRequest = 0;
        v3Copy(Delta_H_B, s->deltaH.torqueRequestBody);
        msg_stamp(&s->deltaH.h, now);
    }
}
/* thruster geometry of AP:73-156 and MOOG Monarc-1 values of [BSK: simIncludeThruster.py] */
static const double THR_LOC[8][3] = {
    {3.874945160902288e-2, -1.206182747348013, 0.85245}, {3.874945160902288e-2, -1.206182747348013, -0.85245},
    {-3.8749451609022656e-2, -1.206182747348013, 0.85245}, {-3.8749451609022656e-2, -1.206182747348013, -0.85245},
    {-3.874945160902288e-2, 1.206182747348013, 0.85245}, {-3.874945160902288e-2, 1.206182747348013, -0.85245},
    {3.8749451609022656e-2, 1.206182747348013, 0.85245}, {3.8749451609022656e-2, 1.206182747348013, -0.85245}};
static const double THR_DIR[8][3] = {
    {-0.7071067811865476, 0.7071067811865475, 0.0}, {-0.7071067811865476, 0.7071067811865475, 0.0},
    {0.7071067811865475, 0.7071067811865476, 0.0}, {0.7071067811865475, 0.7071067811865476, 0.0},
    {0.7071067811865476, -0.7071067811865475, 0.0}, {0.7071067811865476, -0.7071067811865475, 0.0},
    {-0.7071067811865475, -0.7071067811865476, 0.0}, {-0.7071067811865475, -0.7071067811865476, 0.0}};
#define THR_MAX_THRUST 0.9
#define THR_MIN_ON_TIME 0.020

static void tfm_findMinimumNormForce(const double C[3][3], int numControlAxes, double epsilon,
                                     double D[3][MAX_THR], const double Lr_B_Bar[3], int numForces, double F[MAX_THR])
{ /* [BSK: thrForceMapping.c findMinimumNormForce] */
    double CD[3][MAX_THR], CDCDT[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, CDCDTInverse[3][3], w[3];
    for (int i = 0; i < MAX_THR; i++) F[i] = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < numForces; j++) {
            CD[i][j] = 0.0;
            for (int k = 0; k < 3; k++) CD[i][j] += C[i][k] * D[k][j];
        }
    for (int i = 0; i < numControlAxes; i++)
        for (int j = 0; j < numControlAxes; j++) {
            CDCDT[i][j] = 0.0;
            for (int k = 0; k < numForces; k++) CDCDT[i][j] += CD[i][k] * CD[j][k];
        }
    if (m33Determinant(CDCDT) > epsilon) {
        m33Inverse(CDCDT, CDCDTInverse);
        m33MultV3(CDCDTInverse, Lr_B_Bar, w);
        for (int i = 0; i < numForces; i++) {
            F[i] = 0.0;
            for (int k = 0; k < 3; k++) F[i] += CD[k][i] * w[k];
        }
    }
}
static double tfm_computeTorqueAngErr(double D[3][MAX_THR], const double BLr_B[3], int numForces, double epsilon,
                                      const double F[MAX_THR], const double FMag[MAX_THR])
{ /* [BSK: thrForceMapping.c computeTorqueAngErr] */
    double returnAngle = 0.0;
    if (v3Norm(BLr_B) > epsilon) {
        double tauActual_B[3] = {0, 0, 0}, BLr_hat_B[3], LrEffector_B[3];
        for (int i = 0; i < numForces; i++) {
            double thrusterForce = fabs(F[i]) < FMag[i] ? F[i] : FMag[i] * fabs(F[i]) / F[i];
            LrEffector_B[0] = D[0][i]; LrEffector_B[1] = D[1][i]; LrEffector_B[2] = D[2][i];
            v3Scale(thrusterForce, LrEffector_B, LrEffector_B);
            v3Add(tauActual_B, LrEffector_B, tauActual_B);
        }
        v3Normalize(tauActual_B, tauActual_B);
        v3Normalize(BLr_B, BLr_hat_B);
        if (v3Dot(BLr_hat_B, tauActual_B) < 1.0) returnAngle = safeAcos(v3Dot(BLr_hat_B, tauActual_B));
    }
    return returnAngle;
}
static void tfm_map(const double controlAxes_B[9], int thrForceSign, double epsilon, double angErrThresh, int use2ndLoop,
                    const double Lr_in[3], double F[MAX_THR], double *angErrOut)
{ /* [BSK: thrForceMapping.c Update_thrForceMapping]; CoM_B = 0; the "torque" is Delta H (SIM:464) */
    double D[3][MAX_THR], Dbar[3][MAX_THR], C[3][3], Lr_B[3], Lr_offset[3] = {0, 0, 0}, Lr_B_Bar[3], FMag[MAX_THR];
    double Fbar[MAX_THR];
    int numControlAxes = 0, thrusterUsed[MAX_THR];
    memset(D, 0, sizeof(D)); memset(Dbar, 0, sizeof(Dbar)); memset(C, 0, sizeof(C));
    for (int i = 0; i < 3; i++) if (v3Norm(&controlAxes_B[3 * numControlAxes]) > epsilon) numControlAxes++;
    v3Copy(Lr_in, Lr_B);
    for (int i = 0; i < MAX_THR; i++) {
        double rCrossGt[3], LrLocal[3];
        FMag[i] = THR_MAX_THRUST;
        v3Cross(THR_LOC[i], THR_DIR[i], rCrossGt);
        for (int j = 0; j < 3; j++) D[j][i] = rCrossGt[j];
        if (thrForceSign < 0) { v3Scale(FMag[i], rCrossGt, LrLocal); v3Subtract(Lr_offset, LrLocal, Lr_offset); }
    }
    v3Add(Lr_offset, Lr_B, Lr_B);
    for (int i = 0; i < numControlAxes; i++) v3Copy(&controlAxes_B[3 * i], C[i]);
    m33MultV3(C, Lr_B, Lr_B_Bar);
    tfm_findMinimumNormForce(C, numControlAxes, epsilon, D, Lr_B_Bar, MAX_THR, F);
    if (thrForceSign > 0) { /* substractMin */
        double minValue = 0.0;
        for (int i = 0; i < MAX_THR; i++) if (F[i] < minValue) minValue = F[i];
        for (int i = 0; i < MAX_THR; i++) F[i] -= minValue;
    }
    if (thrForceSign < 0 || use2ndLoop) {
        int counterPosForces = 0, c = 0;
        memset(thrusterUsed, 0, sizeof(thrusterUsed));
        for (int i = 0; i < MAX_THR; i++)
            if (F[i] * thrForceSign > epsilon) {
                thrusterUsed[i] = 1;
                for (int j = 0; j < 3; j++) Dbar[j][counterPosForces] = D[j][i];
                counterPosForces++;
            }
        tfm_findMinimumNormForce(C, numControlAxes, epsilon, Dbar, Lr_B_Bar, counterPosForces, Fbar);
        if (thrForceSign > 0) {
            double minValue = 0.0;
            for (int i = 0; i < counterPosForces; i++) if (Fbar[i] < minValue) minValue = Fbar[i];
            for (int i = 0; i < counterPosForces; i++) Fbar[i] -= minValue;
        }
        for (int i = 0; i < MAX_THR; i++) { if (thrusterUsed[i]) { F[i] = Fbar[c]; c++; } else F[i] = 0.0; }
    }
    double angErr = tfm_computeTorqueAngErr(D, Lr_B_Bar, MAX_THR, epsilon, F, FMag);
    if (angErr > angErrThresh) {
        double maxFractUse = 0.0;
        for (int i = 0; i < MAX_THR; i++)
            if (FMag[i] > 0 && fabs(F[i]) / FMag[i] > maxFractUse) maxFractUse = fabs(F[i]) / FMag[i];
        if (maxFractUse > 1.0) {
            for (int i = 0; i < MAX_THR; i++) F[i] = (1.0 / maxFractUse) * F[i];
            angErr = tfm_computeTorqueAngErr(D, Lr_B_Bar, MAX_THR, epsilon, F, FMag);
        }
    }
    if (angErrOut) *angErrOut = angErr;
}
void orc_thr_force_mapping(const double Lr[3], double F[8], double *angErr)
{
    const double axes[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    tfm_map(axes, 1, 0.0, 0.0, 0, Lr, F, angErr);
}
static void thrForceMapping_update(orc_leo_sim *s, uint64_t now)
{
    double Lr[3] = {0, 0, 0};
    if (msg_written(&s->deltaH.h)) v3Copy(s->deltaH.torqueRequestBody, Lr);
    tfm_map(s->controlAxes_B, s->thrForceSign, s->tfm_epsilon, s->tfm_angErrThresh, s->tfm_use2ndLoop, Lr,
            s->deltaP.thrForce, &s->outTorqAngErr);
    msg_stamp(&s->deltaP.h, now);
}
static void thrMomentumDumping_update(orc_leo_sim *s, uint64_t now)
{ /* [BSK: thrMomentumDumping.c] */
    double tOnOut[MAX_THR] = {0};
    if (s->dumpPriorTime != 0) {
        double dt = (double)(now - s->dumpPriorTime) * NANO2SEC;
        if (dt < 0.0) dt = 0.0;
        uint64_t timeOfDeltaHMsg = msg_written(&s->deltaH.h) ? s->deltaH.h.write_ns : 0;
        if (s->lastDeltaHInMsgTime != timeOfDeltaHMsg) {
            s->lastDeltaHInMsgTime = timeOfDeltaHMsg;
            s->thrDumpingCounter = 0;
            for (int i = 0; i < s->nThr; i++) s->thrOnTimeRemaining[i] = s->deltaP.thrForce[i] / s->thr[i].MaxThrust;
        }
        if (s->thrDumpingCounter <= 0) {
            for (int i = 0; i < s->nThr; i++) tOnOut[i] = s->thrOnTimeRemaining[i];
            for (int i = 0; i < s->nThr; i++) if (s->thrOnTimeRemaining[i] > 0.0) s->thrOnTimeRemaining[i] -= dt;
            s->thrDumpingCounter = s->maxCounterValue;
        } else {
            s->thrDumpingCounter -= 1;
        }
        for (int i = 0; i < s->nThr; i++) {
            if (tOnOut[i] < s->thrMinFireTime) tOnOut[i] = 0.0;
            if (s->thrOnTimeRemaining[i] < 0.0) s->thrOnTimeRemaining[i] = 0.0;
            if (tOnOut[i] >= dt) tOnOut[i] = dt;
        }
    }
    s->dumpPriorTime = now;
    for (int i = 0; i < MAX_THR; i++) s->thrOnTime.OnTimeRequest[i] = tOnOut[i];
    msg_stamp(&s->thrOnTime.h, now);
}
static void thrMomentumManagement_reset(orc_leo_sim *s) { s->initRequest = 1; }
static void thrMomentumDumping_reset(orc_leo_sim *s)
{
    s->dumpPriorTime = 0;
    s->thrDumpingCounter = 0;
    memset(s->thrOnTimeRemaining, 0, sizeof(s->thrOnTimeRemaining));
    s->lastDeltaHInMsgTime = 0;
}

/* ================================ scenario wiring (SIM:67-533) ================================= */
void orc_leo_default_cfg(orc_leo_cfg *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->dynRate = 0.1; cfg->fswRate = 1.0; cfg->step_duration = 180.;
}
orc_leo_sim *orc_leo_create(const orc_leo_ic *ic, const orc_leo_cfg *cfg)
{
    orc_leo_sim *s = (orc_leo_sim *)calloc(1, sizeof(orc_leo_sim));
    s->cfg = *cfg; s->ic = *ic;
    s->scenario = 0; s->mu_central = MU_EARTH;
    /* tasks, in creation order with their priorities (SIM:101-103, 383-386) */
    sched_add_task(s, T_DYN, cfg->dynRate, -1);
    sched_add_task(s, T_SPICE, cfg->step_duration, -1);
    sched_add_task(s, T_ENV, cfg->dynRate, -1);
    sched_add_task(s, T_SUNPOINT, cfg->fswRate, 100);
    sched_add_task(s, T_NADIRPOINT, cfg->fswRate, 100);
    sched_add_task(s, T_MRPCONTROL, cfg->fswRate, 50);
    sched_add_task(s, T_RWDESAT, cfg->fswRate, 100);
    /* set_dynamics (SIM:195-368) */
    double mass = 330, width = 1.38, depth = 1.04, height = 1.58;
    s->mHub = mass;
    s->IHub[0][0] = 1. / 12. * mass * (pow(width, 2.) + pow(depth, 2.));
    s->IHub[1][1] = 1. / 12. * mass * (pow(depth, 2.) + pow(height, 2.));
    s->IHub[2][2] = 1. / 12. * mass * (pow(width, 2.) + pow(height, 2.));
    v3Copy(ic->rN, s->r); v3Copy(ic->vN, s->v); v3Copy(ic->sigma_init, s->sigma); v3Copy(ic->omega_init, s->omega);
    s->planetRadius = REQ_EARTH_KM * 1000.; s->baseDensity = 1.22; s->scaleHeight = 8e3;
    { /* facets SIM:274-281 */
        const double A[8] = {0.2 * 0.3, 0.2 * 0.3, 0.1 * 0.2, 0.1 * 0.2, 0.1 * 0.3, 0.1 * 0.3, 1. * 2., 1. * 2.};
        const double N[8][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}};
        const double Lc[8][3] = {{0.05, 0, 0}, {0.05, 0, 0}, {0, 0.15, 0}, {0, -0.15, 0}, {0, 0, 0.1}, {0, 0, -0.1}, {0, 2., 0}, {0, 2., 0}};
        s->nFacet = 8;
        for (int i = 0; i < 8; i++) { s->facetArea[i] = A[i]; s->facetCd[i] = 2.2; v3Copy(N[i], s->facetN[i]); v3Copy(Lc[i], s->facetLoc[i]); }
    }
    v3Scale(2e-4, ic->disturbance_vector, s->extTorquePntB_B);           /* SIM:295 */
    { /* balancedHR16Triad AP:20-37 + [BSK: simIncludeRW.py Honeywell_HR16, maxMomentum=50]; or the opNav pyramid
         (BSK_OpNavDynamics.py:269-293): gsHat = Mi(-az,3) Mi(el,2) [1,0,0] */
        const double gs3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        s->nRW = cfg->rw_set == 1 ? 4 : 3;
        for (int i = 0; i < s->nRW; i++) {
            RWheel *w = &s->rw[i];
            if (cfg->rw_set == 1) {
                double el = 40.0 * PI_D / 180.0, az = (45.0 + 90.0 * i) * PI_D / 180.0;
                w->gsHat_B[0] = cos(el) * cos(az); w->gsHat_B[1] = cos(el) * sin(az); w->gsHat_B[2] = sin(el);
            } else {
                v3Copy(gs3[i], w->gsHat_B);
            }
            w->Omega_max = 6000.0 * RPM; w->u_max = 0.200; w->u_min = 0.0;  /* useMinTorque False */
            w->Js = 50. / w->Omega_max;
            double rpm = i < 3 ? ic->wheelSpeeds_rpm[i]
                               : (ic->wheelSpeeds_rpm[0] + ic->wheelSpeeds_rpm[1] + ic->wheelSpeeds_rpm[2]) / 3.0;
            w->Omega = rpm * RPM;                                         /* SIM:303-305 */
        }
    }
    s->nThr = 8;
    for (int i = 0; i < 8; i++) {
        v3Copy(THR_LOC[i], s->thr[i].thrLoc_B); v3Copy(THR_DIR[i], s->thr[i].thrDir_B);
        s->thr[i].MaxThrust = THR_MAX_THRUST; s->thr[i].MinOnTime = THR_MIN_ON_TIME;
    }
    s->nHat_B[0] = 0; s->nHat_B[1] = -1; s->nHat_B[2] = 0; s->panelArea = 0.2 * 0.3; s->panelEfficiency = 0.20;
    s->nodePowerOut = -5.0;
    s->storageCapacity = 20.0 * 3600.; s->storedCharge = ic->storedCharge_Init;
    /* initial obs SIM:347-351 */
    s->obs[0] = v3Norm(ic->sigma_init); s->obs[1] = v3Norm(ic->omega_init);
    { double w[3], w4 = 0.0; for (int i = 0; i < 3; i++) w[i] = ic->wheelSpeeds_rpm[i];
      if (s->nRW == 4) w4 = (w[0] + w[1] + w[2]) / 3.0;
      s->obs[2] = sqrt(v3Dot(w, w) + w4 * w4); } /* RPM, un-converted (SIM:306,350) */
    s->obs[3] = ic->storedCharge_Init / 3600.0; s->obs[4] = 0.0;
    /* model -> task assignment, in AddModelToTask order (SIM:356-366) */
    task_add_model(s, T_DYN, sc_update);
    task_add_model(s, T_SPICE, spice_update);
    task_add_model(s, T_DYN, atmosphere_update);
    task_add_model(s, T_DYN, drag_update);
    task_add_model(s, T_DYN, simple_nav_update);
    task_add_model(s, T_DYN, rw_update);
    task_add_model(s, T_DYN, thruster_update);
    task_add_model(s, T_ENV, eclipse_update);
    task_add_model(s, T_ENV, solar_panel_update);
    task_add_model(s, T_ENV, battery_update);
    task_add_model(s, T_ENV, power_sink_update);
    /* set_fsw (SIM:371-490) */
    memcpy(s->ISC_fsw, s->IHub, sizeof(s->IHub));
    s->sigma_R0N[0] = 1; s->sigma_R0N[1] = 0; s->sigma_R0N[2] = 0;
    { const double ax[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; memcpy(s->controlAxes_B, ax, sizeof(ax)); }
    s->K = 7; s->Ki = -1.0; s->P = 35;
    s->hs_min = 4.; s->thrForceSign = 1; s->maxCounterValue = 4; s->thrMinFireTime = 0.002;
    s->tfm_epsilon = 0.0; s->tfm_angErrThresh = 0.0; s->tfm_use2ndLoop = 0; /* zero-initialised config struct */
    task_add_model(s, T_SUNPOINT, inertial3D_update);
    task_add_model(s, T_NADIRPOINT, hillPoint_update);
    task_add_model(s, T_MRPCONTROL, MRP_Feedback_update);       /* quirk Q1: controller BEFORE tracking error */
    task_add_model(s, T_MRPCONTROL, attTrackingError_update);
    task_add_model(s, T_MRPCONTROL, rwMotorTorque_update);
    task_add_model(s, T_RWDESAT, thrMomentumManagement_update);
    task_add_model(s, T_RWDESAT, thrForceMapping_update);
    task_add_model(s, T_RWDESAT, thrMomentumDumping_update);
    /* set_logging (SIM:493-533) */
    uint64_t samplingTime = sec2nano(cfg->step_duration);
    log_add(s, &s->scState.h, s->scState.r_BN_N, 12 * sizeof(double), samplingTime);
    log_add(s, &s->navAtt.h, s->navAtt.sigma_BN, 9 * sizeof(double), samplingTime);
    log_add(s, &s->rwSpeeds.h, s->rwSpeeds.wheelSpeeds, MAX_RW * sizeof(double), samplingTime);
    log_add(s, &s->attRef.h, s->attRef.sigma_RN, 9 * sizeof(double), samplingTime);
    log_add(s, &s->attGuid.h, s->attGuid.sigma_BR, 12 * sizeof(double), samplingTime);
    log_add(s, &s->battery.h, &s->battery.storageCapacity, 3 * sizeof(double), samplingTime);
    log_add(s, &s->eclipseMsg.h, &s->eclipseMsg.shadowFactor, sizeof(double), samplingTime);
    /* InitializeSimulationAndDiscover (SIM:115): SelfInit/CrossInit/Reset(0) of every model */
    spice_update(s, 0);                 /* SpiceInterface::Reset writes the planet states */
    thrMomentumManagement_reset(s);
    thrMomentumDumping_reset(s);
    s->mrpPriorTime = 0;
    s->batPreviousTime = 0;
    s->thrPrevCommandTime = 0xFFFFFFFFFFFFFFFFull;   /* ThrusterDynamicEffector ctor */
    s->next_task_time = 0;
    return s;
}
void orc_leo_destroy(orc_leo_sim *s) { free(s); }
void orc_leo_initial_obs(const orc_leo_sim *s, double obs[5]) { memcpy(obs, s->obs, 5 * sizeof(double)); }

int orc_leo_run_sim(orc_leo_sim *s, int action, double obs[5])
{ /* SIM:535-644.  modeRequest = str(action): only "0", "1", "2" switch modes */
    int sim_over = 0;
    if (action == 0) {
        for (int i = 0; i < N_TASKS; i++) s->tasks[i].active = 1;      /* enableAllTasks */
        s->tasks[T_SUNPOINT].active = 0; s->tasks[T_RWDESAT].active = 0;
        s->tasks[T_NADIRPOINT].active = 1; s->tasks[T_MRPCONTROL].active = 1;
    } else if (action == 1) {
        for (int i = 0; i < N_TASKS; i++) s->tasks[i].active = 1;
        s->tasks[T_NADIRPOINT].active = 0; s->tasks[T_RWDESAT].active = 0;
        s->tasks[T_SUNPOINT].active = 1; s->tasks[T_MRPCONTROL].active = 1;
    } else if (action == 2) {
        for (int i = 0; i < N_TASKS; i++) s->tasks[i].active = 1;
        thrMomentumManagement_reset(s);                                 /* SIM:580 */
        thrMomentumDumping_reset(s);                                    /* SIM:581 */
        s->tasks[T_NADIRPOINT].active = 0; s->tasks[T_SUNPOINT].active = 0;
        s->tasks[T_SUNPOINT].active = 1; s->tasks[T_MRPCONTROL].active = 1; s->tasks[T_RWDESAT].active = 1;
    }
    s->simTime += s->cfg.step_duration;
    sim_step_until_stop(s, sec2nano(s->simTime));
    /* pullMultiMessageLogData(..., numRecords=1): last logged record of each message (SIM:598-637) */
    const double *scRec = (const double *)s->logs[0].last;      /* r_BN_N at [0..2] */
    const double *navRec = (const double *)s->logs[1].last;     /* omega_BN_B at [3..5] */
    const double *rwRec = (const double *)s->logs[2].last;
    const double *guidRec = (const double *)s->logs[4].last;    /* sigma_BR at [0..2] */
    const double *batRec = (const double *)s->logs[5].last;     /* storageLevel at [1] */
    const double *eclRec = (const double *)s->logs[6].last;
    s->obs[0] = v3Norm(guidRec);
    s->obs[1] = v3Norm(navRec + 3);
    { double w2 = 0.0; for (int i = 0; i < s->nRW; i++) w2 += rwRec[i] * rwRec[i]; s->obs[2] = sqrt(w2); } /* wheelSpeeds[0:3] (SIM:636); all four in the stress config */
    s->obs[3] = batRec[1] / 3600.;
    s->obs[4] = eclRec[0];
    if (v3Norm(scRec) < (REQ_EARTH_KM / 1000.)) sim_over = 1;           /* quirk Q6 */
    memcpy(obs, s->obs, 5 * sizeof(double));
    return sim_over;
}
void orc_leo_get_state(const orc_leo_sim *s, orc_leo_state *o)
{
    memset(o, 0, sizeof(*o));
    v3Copy(s->r, o->r_BN_N); v3Copy(s->v, o->v_BN_N); v3Copy(s->sigma, o->sigma_BN); v3Copy(s->omega, o->omega_BN_B);
    for (int i = 0; i < s->nRW; i++) { o->Omega[i] = s->rw[i].Omega; o->u_current[i] = s->rw[i].u_current; }
    o->storedCharge = s->storedCharge; o->shadowFactor = s->eclipseMsg.shadowFactor; o->density = s->dragDensity;
    v3Copy(s->attGuid.sigma_BR, o->sigma_BR); v3Copy(s->attGuid.omega_BR_B, o->omega_BR_B); v3Copy(s->attRef.sigma_RN, o->sigma_RN);
    v3Copy(s->cmdTorque.torqueRequestBody, o->Lr);
    for (int i = 0; i < MAX_THR; i++) {
        o->thrOnCmd[i] = s->thr[i].ThrustOnCmd; o->thrOnTimeRemaining[i] = s->thrOnTimeRemaining[i];
        o->thr_fire_count[i] = s->thr[i].fireCounter;
        if (s->thr[i].ThrustFactor > 0.0) o->thr_factor_mask |= 1 << i;
    }
    v3Copy(s->deltaH.torqueRequestBody, o->deltaH);
    v3Copy(s->sunMsg.PositionVector, o->sun_r); v3Copy(s->sunMsg.VelocityVector, o->sun_v);
    o->mrp_switch_count = s->MRPSwitchCount;
    o->dump_counter = s->thrDumpingCounter; o->init_request = s->initRequest;
    o->task_mask = (s->tasks[T_SUNPOINT].active ? 1 : 0) | (s->tasks[T_NADIRPOINT].active ? 2 : 0) |
                   (s->tasks[T_MRPCONTROL].active ? 4 : 0) | (s->tasks[T_RWDESAT].active ? 8 : 0);
    o->sim_nanos = s->current_nanos;
}

/* ================================ gym layer (ENV:20-216) ======================================= */
struct orc_leo_env {
    orc_leo_cfg cfg;
    orc_leo_sim *sim;
    int max_length, curr_step, episode_over;
    double wheel_limit, power_max, reward_mult, failure_penalty, reward_total;
};
orc_leo_env *orc_env_create(const orc_leo_cfg *cfg)
{
    orc_leo_env *e = (orc_leo_env *)calloc(1, sizeof(*e));
    e->cfg = *cfg;
    e->max_length = 3 * 180;               /* ENV:25 */
    e->wheel_limit = 3000 * RPM;           /* ENV:36 */
    e->power_max = 20.0;                   /* ENV:37 */
    e->reward_mult = 1. / e->max_length;   /* ENV:41 */
    e->failure_penalty = 1;                /* ENV:42 */
    return e;
}
void orc_env_destroy(orc_leo_env *e) { if (e->sim) orc_leo_destroy(e->sim); free(e); }
orc_leo_sim *orc_env_sim(orc_leo_env *e) { return e->sim; }
/* the env attributes a user of the reference may change before reset (ENV:25, :41: reward_mult = 1/max_length) */
void orc_env_set_max_length(orc_leo_env *e, int max_length) { e->max_length = max_length; e->reward_mult = 1. / max_length; }
/* ENV:130-136: the episode record `info['episode'] = {'r': self.reward_total, 'l': self.curr_step}` is assembled BEFORE
 * `self.curr_step += 1` (ENV:144); orc_env_step has already incremented, hence the - 1 */
void orc_env_episode(const orc_leo_env *e, double *r, int *l) { *r = e->reward_total; *l = e->curr_step - 1; }
void orc_env_reset(orc_leo_env *e, const orc_leo_ic *ic, double ob[5])
{ /* ENV:172-191 / 202-216 */
    e->episode_over = 0; e->curr_step = 0; e->reward_total = 0;
    if (e->sim) orc_leo_destroy(e->sim);
    e->sim = orc_leo_create(ic, &e->cfg);
    orc_leo_initial_obs(e->sim, ob);
    ob[2] = ob[2] / e->wheel_limit;
    ob[3] = ob[3] / e->power_max;
}
void orc_env_step(orc_leo_env *e, int action, orc_env_out *out)
{ /* ENV:65-145 */
    int reason = 0;
    if (e->curr_step >= e->max_length) { e->episode_over = 1; reason |= 1; }      /* ENV:98-99, quirk Q9 */
    double obs[5];
    int sim_over = orc_leo_run_sim(e->sim, action, obs);
    double reward = 0;
    if (action == 0) reward = fabs(e->reward_mult / (1. + pow(obs[0], 2.0)));     /* ENV:168-169 */
    e->reward_total += reward;
    obs[2] = obs[2] / e->wheel_limit;
    obs[3] = obs[3] / e->power_max;
    if (obs[2] > 1) { e->episode_over = 1; reward -= e->failure_penalty; e->reward_total -= e->failure_penalty; reason |= 2; }
    if (obs[3] == 0) { e->episode_over = 1; reward -= e->failure_penalty; e->reward_total -= e->failure_penalty; reason |= 4; }
    if (sim_over) { e->episode_over = 1; reason |= 8; }
    e->curr_step += 1;
    memcpy(out->ob, obs, sizeof(obs));
    out->reward = reward; out->done = e->episode_over; out->reason = reason;
}
void orc_env_step_batch(orc_leo_env **envs, int n, const int *actions, orc_env_out *outs, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < n; i++) orc_env_step(envs[i], actions[i], &outs[i]);
    (void)nthreads;
}
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
