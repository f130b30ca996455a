/*
 * opnav_oracle.c -- scalar FP64 CPU restatement of the opNav environment step
 * (dynamics half + synthetic nav measurement + relativeODuKF), see opnav_oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (Basilisk absent; no golden outputs in the reference).
 *
 * Structure mirrors Basilisk 1.x: one Update function per module, message structs with a "written"
 * header, the two-process order of BSK_masters.py (DynamicsProcess priority 100 runs before FSWProcess
 * priority 10 at every time step, ONM:53-65), models inside a task by descending priority
 * (OND:100-110, ONF:108-164).  It is deliberately NOT shaped like the fused CUDA kernel: the filter uses
 * Householder QR + the Gill-Golub-Murray-Saunders rank-one modification (Basilisk's ukfQRDJustR /
 * ukfCholDownDate), whereas the kernel uses Givens / hyperbolic rank-one sweeps.
 *
 * Confidence tags: [H] textbook form, [M] structure sure / a detail may differ, [L] reconstructed.
 */
#include "opnav_oracle.h"
#include "bsk_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NANO2SEC 1e-9
#define PI_D 3.14159265358979323846
#define NRW 4
#define NCSS 8
#define NST 6
#define NSP 13

void orc_sun_from_mars(double t, double r[3], double v[3], double *j2000_et); /* bsk_oracle.c */
int orc_eph_eval(int kind, double t, double val[3], double rate[3]);          /* bsk_oracle.c */

static const double MU_MARS_DYN = 4.2828371901284001E+13;  /* OND:386 */
static const double MU_MARS_FSW = 42828.314 * 1E9;         /* [BSK: astroConstants.h MU_MARS]*1e9; also ONF:507 */
static const double REQ_MARS_KM = 3396.19;                 /* [BSK: astroConstants.h REQ_MARS] */
static const double RPM = 0.10471975511965977;
static const double D2R = PI_D / 180.0;

/* ---- linear algebra ---- */
static double v3Dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3Norm(const double a[3]) { return sqrt(v3Dot(a, a)); }
static void v3Copy(const double a[3], double r[3]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static void v3SetZero(double r[3]) { r[0] = r[1] = r[2] = 0.0; }
static void v3Scale(double s, const double a[3], double r[3]) { r[0] = s * a[0]; r[1] = s * a[1]; r[2] = s * a[2]; }
static void v3Add(const double a[3], const double b[3], double r[3]) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static void v3Subtract(const double a[3], const double b[3], double r[3]) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void v3Cross(const double a[3], const double b[3], double r[3])
{
    double t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    v3Copy(t, r);
}
static void v3Normalize(const double a[3], double r[3])
{
    double n = v3Norm(a);
    if (n > 1e-30) v3Scale(1. / n, a, r); else v3SetZero(r);
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
static void m33Inverse(double m[3][3], double r[3][3])
{
    double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                 m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    double t[3][3];
    t[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det; t[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det;
    t[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det; t[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
    t[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det; t[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
    t[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det; t[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det;
    t[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
    memcpy(r, t, sizeof(t));
}

/* ================================ counter-based noise streams =================================== */
/* Philox4x32-10 (Salmon et al. 2011), the same generator the product uses for its device-side initial
 * conditions.  counter = (env_lo, env_hi, tick, stream<<16 | block), key = (seed_lo, seed_hi ^ episode).
 * Each call yields four N(0,1) by Box-Muller on 32-bit uniforms.  (Basilisk seeds one std::mt19937 per module
 * from RNGSeed; an independent-stream generator is the batched equivalent -- the noise STATISTICS follow
 * the reference, the sample values cannot.) */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_opnav_normals(uint64_t seed, uint64_t env, uint64_t episode, uint32_t tick, uint32_t stream, uint32_t block,
                       double out[4])
{
    uint32_t x[4];
    philox4x32((uint32_t)env, (uint32_t)(env >> 32), tick, (stream << 16) | block, (uint32_t)seed,
               (uint32_t)(seed >> 32) ^ (uint32_t)episode, x);
    for (int p = 0; p < 2; p++) {
        double u1 = ((double)x[2 * p] + 1.0) * (1.0 / 4294967296.0);
        double u2 = (double)x[2 * p + 1] * (1.0 / 4294967296.0);
        double rr = sqrt(-2.0 * log(u1)), th = 2.0 * PI_D * u2;
        out[2 * p] = rr * cos(th);
        out[2 * p + 1] = rr * sin(th);
    }
}

/* ================================ SR-UKF utilities ([BSK: fswAlgorithms/attDetermination/_GeneralModuleFiles/ukfUtilities.c]) */
void orc_ukf_qr_just_r(const double *A, int nRow, int nCol, double *R)
{ /* ukfQRDJustR [M]: Householder reflections, only R kept (nCol x nCol, upper triangular) */
    double *M = (double *)malloc(sizeof(double) * (size_t)nRow * (size_t)nCol);
    memcpy(M, A, sizeof(double) * (size_t)nRow * (size_t)nCol);
    for (int k = 0; k < nCol; k++) {
        double nrm = 0.0;
        for (int i = k; i < nRow; i++) nrm += M[i * nCol + k] * M[i * nCol + k];
        nrm = sqrt(nrm);
        if (nrm < 1e-300) continue;
        double alpha = M[k * nCol + k] > 0 ? -nrm : nrm;
        double v0 = M[k * nCol + k] - alpha;
        double vtv = v0 * v0;
        for (int i = k + 1; i < nRow; i++) vtv += M[i * nCol + k] * M[i * nCol + k];
        for (int j = k + 1; j < nCol; j++) {
            double s = v0 * M[k * nCol + j];
            for (int i = k + 1; i < nRow; i++) s += M[i * nCol + k] * M[i * nCol + j];
            s = 2.0 * s / vtv;
            M[k * nCol + j] -= s * v0;
            for (int i = k + 1; i < nRow; i++) M[i * nCol + j] -= s * M[i * nCol + k];
        }
        M[k * nCol + k] = alpha;
        for (int i = k + 1; i < nRow; i++) M[i * nCol + k] = 0.0;
    }
    for (int i = 0; i < nCol; i++) for (int j = 0; j < nCol; j++) R[i * nCol + j] = j >= i ? M[i * nCol + j] : 0.0;
    free(M);
}
int orc_ukf_chol_downdate(const double *rMat, const double *xVec, double beta, int n, double *rOut)
{ /* ukfCholDownDate [M]: rOut rOut^T = rMat rMat^T + beta x x^T for lower-triangular rMat (Gill, Golub, Murray,
     Saunders 1974, method C1); returns -1 when the modification is not positive definite */
    double wVec[NST], bParam[NST + 1];
    for (int i = 0; i < n; i++) wVec[i] = xVec[i];
    for (int i = 0; i < n * n; i++) rOut[i] = 0.0;
    bParam[0] = 1.0;
    for (int i = 0; i < n; i++) {
        double rEl2 = rMat[i * n + i] * rMat[i * n + i];
        double arg = rEl2 + beta / bParam[i] * wVec[i] * wVec[i];
        if (arg < 0.0) return -1;
        rOut[i * n + i] = sqrt(arg);
        bParam[i + 1] = bParam[i] + beta * wVec[i] * wVec[i] / rEl2;
        for (int j = i + 1; j < n; j++) {
            wVec[j] = wVec[j] - wVec[i] / rMat[i * n + i] * rMat[j * n + i];
            rOut[j * n + i] = rOut[i * n + i] / rMat[i * n + i] * rMat[j * n + i];
            rOut[j * n + i] += rOut[i * n + i] * beta * wVec[i] * wVec[j] / rEl2 / bParam[i + 1];
        }
    }
    return 0;
}
int orc_ukf_chol_decomp(const double *A, int n, double *L)
{ /* ukfCholDecomp [H]: lower Cholesky factor */
    for (int i = 0; i < n * n; i++) L[i] = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            if (i == j) { if (s <= 0.0) return -1; L[i * n + i] = sqrt(s); }
            else L[i * n + j] = s / L[j * n + j];
        }
    return 0;
}
static void lower_inverse(const double *L, int n, double *Li)
{ /* ukfLInv [H]: inverse of a lower-triangular matrix by forward substitution */
    for (int i = 0; i < n * n; i++) Li[i] = 0.0;
    for (int c = 0; c < n; c++)
        for (int i = c; i < n; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= L[i * n + k] * Li[k * n + c];
            Li[i * n + c] = s / L[i * n + i];
        }
}
static void two_body(const double x[NST], double mu, double dx[NST])
{ /* relODuKFTwoBodyDyn [H] */
    double rn = v3Norm(x);
    for (int k = 0; k < 3; k++) { dx[k] = x[3 + k]; dx[3 + k] = -mu / (rn * rn * rn) * x[k]; }
}
void orc_ukf_state_prop(double x[NST], double mu, double dt)
{ /* relODStateProp [H]: classical RK4 over dt */
    double k1[NST], k2[NST], k3[NST], k4[NST], s[NST];
    two_body(x, mu, k1);
    for (int i = 0; i < NST; i++) s[i] = x[i] + dt / 2.0 * k1[i];
    two_body(s, mu, k2);
    for (int i = 0; i < NST; i++) s[i] = x[i] + dt / 2.0 * k2[i];
    two_body(s, mu, k3);
    for (int i = 0; i < NST; i++) s[i] = x[i] + dt * k3[i];
    two_body(s, mu, k4);
    for (int i = 0; i < NST; i++) x[i] = x[i] + dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
}
void orc_ukf_init(orc_ukf *f, const double state[NST], const double covar[NST * NST], const double qNoise[NST * NST],
                  double mu, double noiseSF)
{ /* Reset_relODuKF [M]: weights of the scaled unscented transform, Cholesky factors of P0 and Q */
    memset(f, 0, sizeof(*f));
    const double alpha = 0.02, beta = 2.0, kappa = 0.0;            /* ONF:499-501 */
    double lambda = alpha * alpha * (NST + kappa) - NST;
    f->gamma = sqrt(NST + lambda);
    f->wM[0] = lambda / (NST + lambda);
    f->wC[0] = lambda / (NST + lambda) + (1 - alpha * alpha + beta);
    for (int i = 1; i < NSP; i++) { f->wM[i] = 1.0 / 2.0 * 1.0 / (NST + lambda); f->wC[i] = f->wM[i]; }
    memcpy(f->state, state, sizeof(f->state));
    memcpy(f->covar, covar, sizeof(f->covar));
    orc_ukf_chol_decomp(covar, NST, f->sBar);
    double sq[NST * NST];
    orc_ukf_chol_decomp(qNoise, NST, sq);
    for (int i = 0; i < NST; i++) for (int j = 0; j < NST; j++) f->sQnoise[i * NST + j] = sq[j * NST + i];
    f->mu = mu; f->noiseSF = noiseSF; f->timeTag = 0.0;
}
void orc_ukf_time_update(orc_ukf *f, double updateTime)
{ /* relODuKFTimeUpdate [M] */
    double dt = updateTime - f->timeTag;
    double procNoise[NST * NST], AT[(2 * NST + NST) * NST], rAT[NST * NST], sBarNew[NST * NST], sBarUp[NST * NST], xErr[NST];
    double SP[NSP * NST], xBar[NST];
    memcpy(procNoise, f->sQnoise, sizeof(procNoise));
    memcpy(&SP[0], f->state, sizeof(double) * NST);
    orc_ukf_state_prop(&SP[0], f->mu, dt);
    for (int k = 0; k < NST; k++) xBar[k] = f->wM[0] * SP[k];
    for (int i = 0; i < NST; i++) {
        for (int sgn = 0; sgn < 2; sgn++) {
            int Index = i + 1 + sgn * NST;
            double *sp = &SP[Index * NST];
            for (int k = 0; k < NST; k++) sp[k] = f->state[k] + (sgn ? -f->gamma : f->gamma) * f->sBar[k * NST + i]; /* column i */
            orc_ukf_state_prop(sp, f->mu, dt);
            for (int k = 0; k < NST; k++) xBar[k] += f->wM[Index] * sp[k];
        }
    }
    for (int i = 0; i < 2 * NST; i++)
        for (int k = 0; k < NST; k++) AT[i * NST + k] = sqrt(f->wC[i + 1]) * (SP[(i + 1) * NST + k] - xBar[k]);
    for (int k = 0; k < 3; k++) { /* process noise scaled with the step */
        procNoise[k * NST + k] *= dt * dt / 2;
        procNoise[(k + 3) * NST + (k + 3)] *= dt;
    }
    memcpy(&AT[2 * NST * NST], procNoise, sizeof(procNoise));
    orc_ukf_qr_just_r(AT, 3 * NST, NST, rAT);
    for (int i = 0; i < NST; i++) for (int j = 0; j < NST; j++) sBarNew[i * NST + j] = rAT[j * NST + i];
    for (int k = 0; k < NST; k++) xErr[k] = SP[k] - xBar[k];
    if (orc_ukf_chol_downdate(sBarNew, xErr, f->wC[0], NST, sBarUp) < 0) { f->n_bad++; return; }   /* relODuKFCleanUpdate */
    memcpy(f->sBar, sBarUp, sizeof(sBarUp));
    for (int i = 0; i < NST; i++) for (int j = 0; j < NST; j++) {
        double s = 0; for (int k = 0; k < NST; k++) s += f->sBar[i * NST + k] * f->sBar[j * NST + k];
        f->covar[i * NST + j] = s;
    }
    memcpy(f->SP, SP, sizeof(SP)); memcpy(f->xBar, xBar, sizeof(xBar));
    memcpy(f->state, &SP[0], sizeof(double) * NST);     /* the propagated central point becomes the estimate */
    f->timeTag = updateTime;
}
void orc_ukf_meas_update(orc_ukf *f, const double obs[3], const double covar_N[9])
{ /* relODuKFMeasUpdate + relODuKFMeasModel [M]: y = first three states, R = noiseSF * covar_N */
    double yMeas[NSP * 3], yBar[3] = {0, 0, 0}, AT[(2 * NST + 3) * 3], measNoise[9], qChol[9], rAT[9], sy[9], syUp[9], dy0[3];
    for (int i = 0; i < NSP; i++) for (int j = 0; j < 3; j++) yMeas[i * 3 + j] = f->SP[i * NST + j];
    for (int i = 0; i < 9; i++) measNoise[i] = f->noiseSF * covar_N[i];
    for (int i = 0; i < NSP; i++) for (int j = 0; j < 3; j++) yBar[j] += f->wM[i] * yMeas[i * 3 + j];
    for (int i = 0; i < 2 * NST; i++) for (int j = 0; j < 3; j++) AT[i * 3 + j] = sqrt(f->wC[i + 1]) * (yMeas[(i + 1) * 3 + j] - yBar[j]);
    if (orc_ukf_chol_decomp(measNoise, 3, qChol) < 0) { f->n_bad++; return; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) AT[(2 * NST + i) * 3 + j] = qChol[j * 3 + i];   /* L^T so that AT^T AT adds R */
    orc_ukf_qr_just_r(AT, 2 * NST + 3, 3, rAT);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sy[i * 3 + j] = rAT[j * 3 + i];
    for (int j = 0; j < 3; j++) dy0[j] = yMeas[j] - yBar[j];
    if (orc_ukf_chol_downdate(sy, dy0, f->wC[0], 3, syUp) < 0) { f->n_bad++; return; }
    memcpy(sy, syUp, sizeof(sy));
    double pXY[NST * 3];
    memset(pXY, 0, sizeof(pXY));
    for (int i = 0; i < NSP; i++)
        for (int a = 0; a < NST; a++) for (int b = 0; b < 3; b++)
            pXY[a * 3 + b] += f->wC[i] * (f->SP[i * NST + a] - f->xBar[a]) * (yMeas[i * 3 + b] - yBar[b]);
    /* K = Pxy (Sy Sy^T)^-1 = Pxy Sy^-T Sy^-1 */
    double syInv[9], kMat[NST * 3], tmp[NST * 3];
    lower_inverse(sy, 3, syInv);
    for (int a = 0; a < NST; a++) for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < 3; k++) s += pXY[a * 3 + k] * syInv[b * 3 + k]; tmp[a * 3 + b] = s; }
    for (int a = 0; a < NST; a++) for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < 3; k++) s += tmp[a * 3 + k] * syInv[k * 3 + b]; kMat[a * 3 + b] = s; }
    double innov[3];
    for (int j = 0; j < 3; j++) innov[j] = obs[j] - yBar[j];
    double newState[NST];
    for (int a = 0; a < NST; a++) newState[a] = f->state[a] + kMat[a * 3] * innov[0] + kMat[a * 3 + 1] * innov[1] + kMat[a * 3 + 2] * innov[2];
    /* U = K Sy; one rank-one down-date of sBar per column of U */
    double sBar[NST * NST], sNext[NST * NST];
    memcpy(sBar, f->sBar, sizeof(sBar));
    for (int c = 0; c < 3; c++) {
        double Ucol[NST];
        for (int a = 0; a < NST; a++) { double s = 0; for (int k = 0; k < 3; k++) s += kMat[a * 3 + k] * sy[k * 3 + c]; Ucol[a] = s; }
        if (orc_ukf_chol_downdate(sBar, Ucol, -1.0, NST, sNext) < 0) { f->n_bad++; return; }
        memcpy(sBar, sNext, sizeof(sBar));
    }
    memcpy(f->sBar, sBar, sizeof(sBar));
    memcpy(f->state, newState, sizeof(newState));
    for (int i = 0; i < NST; i++) for (int j = 0; j < NST; j++) {
        double s = 0; for (int k = 0; k < NST; k++) s += f->sBar[i * NST + k] * f->sBar[j * NST + k];
        f->covar[i * NST + j] = s;
    }
}

/* ================================ synthetic camera + circle finder + pixelLineConverter ========= */
#define CAM_RES 512.0                    /* OND:138 */
#define CAM_FOV (55.0 * PI_D / 180.0)    /* OND:141 */
#define HOUGH_MIN_RADIUS 20.0            /* ONF:463 */
int orc_opnav_project_circle(const double r_C[3], double R, double c[3])
{ /* pinhole image of the planet: centre of the disc and apparent radius, the exact inverse of pixel_line below.
     Stand-in for camera (OND:112-143) -> Vizard -> houghCircles (ONF:452-468). */
    double pX = 2. * tan(CAM_FOV * CAM_RES / CAM_RES / 2.0), pY = 2. * tan(CAM_FOV / 2.0);
    double X = pX / CAM_RES, Y = pY / CAM_RES;
    double d = v3Norm(r_C);
    c[0] = c[1] = c[2] = 0.0;
    if (!(r_C[2] > 0.0) || !(d > R)) return 0;
    c[0] = (r_C[0] / r_C[2] + pX / 2) / X;
    c[1] = (r_C[1] / r_C[2] + pY / 2) / Y;
    c[2] = R / sqrt(d * d - R * R) / X;
    if (c[0] < 0.0 || c[0] >= CAM_RES || c[1] < 0.0 || c[1] >= CAM_RES || c[2] < HOUGH_MIN_RADIUS) return 0;
    return 1;
}
void orc_opnav_pixel_line(const double c[3], double unc, double dcm_CN[3][3], double r_BN_N[3], double covar_N[9])
{ /* [BSK: fswAlgorithms/imageProcessing/pixelLineConverter/pixelLineConverter.c] [M], planetTarget = 2 (Mars, ONF:474) */
    double pX = 2. * tan(CAM_FOV * CAM_RES / CAM_RES / 2.0), pY = 2. * tan(CAM_FOV / 2.0);
    double X = pX / CAM_RES, Y = pY / CAM_RES;
    double rtilde_C[3] = {X * c[0] - pX / 2, Y * c[1] - pY / 2, 1.0}, rHat_BN_C[3], rHat_BN_N[3];
    v3Normalize(rtilde_C, rHat_BN_C);
    v3Scale(-1, rHat_BN_C, rHat_BN_C);
    m33tMultV3(dcm_CN, rHat_BN_C, rHat_BN_N);
    double planetRad = REQ_MARS_KM;                      /* km */
    double denom = sin(atan(X * c[2]));
    double rNorm = planetRad / denom;
    v3Scale(rNorm * 1E3, rHat_BN_N, r_BN_N);
    double x_map = planetRad / denom * X, y_map = planetRad / denom * Y;
    double rho_map = planetRad * (X / sqrt(1 + pow(c[2] * X, 2)) - 1.0 / X * sqrt(1 + pow(c[2] * X, 2)) / pow(c[2], 2));
    double dC[3] = {x_map * x_map * unc, y_map * y_map * unc, rho_map * rho_map * unc};   /* covar_map C_in covar_map^T, diagonal */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += dcm_CN[k][i] * dC[k] * dcm_CN[k][j];
        covar_N[i * 3 + j] = s * 1E6;                    /* km^2 -> m^2 */
    }
}

/* ================================ the simulator ================================================= */
typedef struct { int written; } Hdr;
typedef struct { Hdr h; double r[3], v[3], sigma[3], omega[3]; } StateMsg;
typedef struct { Hdr h; double pos[3], vel[3]; } PlanetMsg;
typedef struct { Hdr h; double sigma[3], omega[3], sun_B[3]; } NavAttMsg;
typedef struct { Hdr h; double r[3], v[3]; } NavTransMsg;
typedef struct { Hdr h; double sigma_RN[3], omega_RN_N[3], domega_RN_N[3]; } AttRefMsg;
typedef struct { Hdr h; double sigma_BR[3], omega_BR_B[3], omega_RN_B[3], domega_RN_B[3]; } AttGuidMsg;
typedef struct { Hdr h; int valid; uint64_t timeTag; double r[3], sigma[3]; } ImageMsg;      /* what the renderer would be shown */
typedef struct { Hdr h; int valid; uint64_t timeTag; double c[3], unc; } CirclesMsg;
typedef struct { Hdr h; int valid; uint64_t timeTag; double r_BN_N[3], covar_N[9]; } OpNavMsg;

struct orc_opnav_sim {
    orc_opnav_cfg cfg; orc_opnav_ic ic;
    uint64_t env_index, episode;
    uint64_t dyn_ns, next_ns;          /* integer-nanosecond clock; next_ns = time of the next task execution */
    double simTime; int modeCounter, first_run;
    /* truth states */
    double mHub, IHub[3][3];
    double r[3], v[3], sigma[3], omega[3], Omega[NRW], u_current[NRW];
    double gs[NRW][3], Js, u_max, Omega_max;
    double timePrevious; int64_t MRPSwitchCount;
    /* messages */
    StateMsg scState; PlanetMsg sunMsg;
    struct { Hdr h; double shadow; } eclipseMsg;
    struct { Hdr h; double y[NCSS]; } cssMsg;
    struct { Hdr h; double Omega[NRW]; } rwSpeeds;
    struct { Hdr h; double u[NRW]; } rwCmd;
    struct { Hdr h; double Lr[3]; } cmdTorque;
    NavAttMsg navAtt, sunPoint; NavTransMsg navTrans; AttRefMsg attRef; AttGuidMsg attGuid;
    ImageMsg image; CirclesMsg circles; OpNavMsg opnav;
    /* simple_nav */
    double navErrors[18]; uint64_t navPrevTime;
    /* camera */
    int cameraIsOn; int64_t n_images;
    /* FSW */
    int t_opNavPointCheat, t_sunSafe, t_mrpRW, t_opNavOD;
    double sigma_R0R[3], K, P, Ki, ISC[3][3];
    double cssN[NCSS][3], sHatBdyCmd[3], eHat180_B[3];
    orc_ukf ukf; int64_t n_meas;
    double obs[4], debug[12];
};

/* ---- DynamicsTask models, in descending priority (OND:100-110) ---- */
static void rw_update(orc_opnav_sim *s)
{ /* ReactionWheelStateEffector::UpdateState (prio 301): latch the motor command, publish the wheel speeds */
    for (int i = 0; i < NRW; i++) {
        double u = s->rwCmd.h.written ? s->rwCmd.u[i] : 0.0;
        if (u > s->u_max) u = s->u_max; else if (u < -s->u_max) u = -s->u_max;
        if (fabs(s->Omega[i]) >= s->Omega_max && s->Omega[i] * u >= 0.0) u = 0.0;
        s->u_current[i] = u;
        s->rwSpeeds.Omega[i] = s->Omega[i];
    }
    s->rwSpeeds.h.written = 1;
}
static void css_update(orc_opnav_sim *s)
{ /* CSSConstellation / CoarseSunSensor::UpdateState (prio 299) [M]: fov 80 deg, scaleFactor 2 (OND:337-338), no noise,
     no albedo, kellyFactor 0; the eclipse message ("eclipse_data_0", OND:360) is the one written at the PREVIOUS tick */
    double r[3] = {0, 0, 0}, sig[3] = {0, 0, 0}, sun[3] = {0, 0, 0}, sc2sun[3], sHat_N[3], sHat_B[3], BN[3][3];
    if (s->scState.h.written) { v3Copy(s->scState.r, r); v3Copy(s->scState.sigma, sig); }
    if (s->sunMsg.h.written) v3Copy(s->sunMsg.pos, sun);
    double shadow = s->eclipseMsg.h.written ? s->eclipseMsg.shadow : 1.0;
    v3Subtract(sun, r, sc2sun);
    v3Normalize(sc2sun, sHat_N);
    orc_MRP2C(sig, BN);
    m33MultV3(BN, sHat_N, sHat_B);
    for (int i = 0; i < NCSS; i++) {
        double d = v3Dot(s->cssN[i], sHat_B);
        double direct = d >= cos(80. * D2R) ? d : 0.0;
        s->cssMsg.y[i] = direct * shadow * 2.0;
    }
    s->cssMsg.h.written = 1;
}
static void eclipse_update(orc_opnav_sim *s)
{ /* Eclipse::UpdateState (prio 204), planet "mars barycenter" at the origin of the zero base (OND:227-230, :401) */
    double zero[3] = {0, 0, 0};
    if (!s->scState.h.written) s->eclipseMsg.shadow = 1.0;   /* degenerate geometry of the unwritten state message */
    else s->eclipseMsg.shadow = orc_eclipse_shadow(s->sunMsg.pos, zero, s->scState.r, REQ_MARS_KM * 1000.);
    s->eclipseMsg.h.written = 1;
}
static void eom(const orc_opnav_sim *s, const double x[16], double dx[16])
{ /* SpacecraftPlus::equationsOfMotion with the balanced-wheel back-substitution; Mars point mass only (OND:382-391);
     thrusters attached but never commanded, extForceTorque zero (OND:232-234) */
    const double *r = x, *v = x + 3, *q = x + 6, *w = x + 9, *Om = x + 12;
    double D[3][3], vecRot[3] = {0, 0, 0};
    memcpy(D, s->IHub, sizeof(D));
    for (int i = 0; i < NRW; i++) {
        double wxg[3];
        v3Cross(w, s->gs[i], wxg);
        for (int a = 0; a < 3; a++) {
            for (int b = 0; b < 3; b++) D[a][b] -= s->Js * s->gs[i][a] * s->gs[i][b];
            vecRot[a] -= s->gs[i][a] * s->u_current[i] + s->Js * Om[i] * wxg[a];
        }
    }
    double Iw[3], wxIw[3], Dinv[3][3], wDot[3];
    m33MultV3((double (*)[3])s->IHub, w, Iw);
    v3Cross(w, Iw, wxIw);
    for (int a = 0; a < 3; a++) vecRot[a] -= wxIw[a];
    m33Inverse(D, Dinv);
    m33MultV3(Dinv, vecRot, wDot);
    double rn = v3Norm(r);
    for (int k = 0; k < 3; k++) { dx[k] = v[k]; dx[3 + k] = -r[k] * MU_MARS_DYN / (rn * rn * rn); dx[9 + k] = wDot[k]; }
    double n2 = v3Dot(q, q), B[3][3], Bw[3];
    B[0][0] = 1 - n2 + 2 * q[0] * q[0]; B[0][1] = 2 * (q[0] * q[1] - q[2]); B[0][2] = 2 * (q[0] * q[2] + q[1]);
    B[1][0] = 2 * (q[1] * q[0] + q[2]); B[1][1] = 1 - n2 + 2 * q[1] * q[1]; B[1][2] = 2 * (q[1] * q[2] - q[0]);
    B[2][0] = 2 * (q[2] * q[0] - q[1]); B[2][1] = 2 * (q[2] * q[1] + q[0]); B[2][2] = 1 - n2 + 2 * q[2] * q[2];
    m33MultV3(B, w, Bw);
    for (int k = 0; k < 3; k++) dx[6 + k] = 0.25 * Bw[k];
    for (int i = 0; i < NRW; i++) dx[12 + i] = s->u_current[i] / s->Js - v3Dot(s->gs[i], wDot);
}
static void sc_update(orc_opnav_sim *s, uint64_t now)
{ /* SpacecraftPlus::UpdateState (prio 201): RK4 over [timePrevious, now], MRP switch, write the state message */
    double newTime = now * NANO2SEC, h = newTime - s->timePrevious;
    double x0[16], x[16], k[16], xo[16];
    memcpy(x0, s->r, 24); memcpy(x0 + 3, s->v, 24); memcpy(x0 + 6, s->sigma, 24); memcpy(x0 + 9, s->omega, 24); memcpy(x0 + 12, s->Omega, 32);
    memcpy(xo, x0, sizeof(xo));
    eom(s, x0, k);
    for (int i = 0; i < 16; i++) { xo[i] += k[i] * (h / 6.0); x[i] = x0[i] + 0.5 * h * k[i]; }
    eom(s, x, k);
    for (int i = 0; i < 16; i++) { xo[i] += k[i] * (h / 3.0); x[i] = x0[i] + 0.5 * h * k[i]; }
    eom(s, x, k);
    for (int i = 0; i < 16; i++) { xo[i] += k[i] * (h / 3.0); x[i] = x0[i] + h * k[i]; }
    eom(s, x, k);
    for (int i = 0; i < 16; i++) xo[i] += k[i] * (h / 6.0);
    memcpy(s->r, xo, 24); memcpy(s->v, xo + 3, 24); memcpy(s->sigma, xo + 6, 24); memcpy(s->omega, xo + 9, 24); memcpy(s->Omega, xo + 12, 32);
    s->timePrevious = newTime;
    if (v3Norm(s->sigma) > 1) {
        double d = v3Dot(s->sigma, s->sigma);
        for (int a = 0; a < 3; a++) s->sigma[a] = -s->sigma[a] / d;
        s->MRPSwitchCount++;
    }
    v3Copy(s->r, s->scState.r); v3Copy(s->v, s->scState.v); v3Copy(s->sigma, s->scState.sigma); v3Copy(s->omega, s->scState.omega);
    s->scState.h.written = 1;
}
static void spice_update(orc_opnav_sim *s, uint64_t now)
{ /* SpiceInterface (prio 200), zeroBase "mars barycenter" (OND:401): analytic stand-in, documented deviation */
    orc_sun_from_mars(now * NANO2SEC, s->sunMsg.pos, s->sunMsg.vel, 0);
    orc_eph_eval(2, now * NANO2SEC, s->sunMsg.pos, s->sunMsg.vel);      /* the table, when one is loaded */
    s->sunMsg.h.written = 1;
}
static const double NAV_P[15] = {10.0, 10.0, 10.0, 0.001, 0.001, 0.001,                       /* OND:238-247 */
                                 1.0 / 36000.0 * PI_D / 180.0, 1.0 / 36000.0 * PI_D / 180.0, 1.0 / 36000.0 * PI_D / 180.0,
                                 0.00005 * PI_D / 180.0, 0.00005 * PI_D / 180.0, 0.00005 * PI_D / 180.0,
                                 0.1 * PI_D / 180.0, 0.1 * PI_D / 180.0, 0.1 * PI_D / 180.0};
static const double NAV_BOUND[15] = {100000.0, 100000.0, 100000.0, 0.1, 0.1, 0.1,               /* OND:248-253 */
                                     1E-18 * PI_D / 180.0, 1E-18 * PI_D / 180.0, 1E-18 * PI_D / 180.0,
                                     1E-18 * PI_D / 180.0, 1E-18 * PI_D / 180.0, 1E-18 * PI_D / 180.0,
                                     5.0 * PI_D / 180.0, 5.0 * PI_D / 180.0, 5.0 * PI_D / 180.0};
static void simple_nav_update(orc_opnav_sim *s, uint64_t now)
{ /* SimpleNav::UpdateState (prio 109) [M]: Gauss-Markov error states (crossTrans: position error integrates the velocity
     error, OND:257), bounded random walk of [BSK: utilities/gauss_markov.cpp], errors applied to the truth */
    if (s->cfg.nav_noise) {
        double dt = (double)(now - s->navPrevTime) * NANO2SEC;
        double ran[16];
        for (int b = 0; b < 4; b++)
            orc_opnav_normals(s->cfg.seed, s->env_index, s->episode, (uint32_t)(now / s->dyn_ns), 1, (uint32_t)b, &ran[4 * b]);
        for (int i = 0; i < 3; i++) s->navErrors[i] += dt * s->navErrors[3 + i];
        for (int i = 0; i < 15; i++) {
            double x = s->navErrors[i], bound = NAV_BOUND[i], rn = ran[i];
            if (bound > 0.0) {
                double stateCalc = fabs(x) > bound * 1E-10 ? fabs(x) : bound;
                double boundCheck = (bound * 2.0 - stateCalc) / stateCalc;
                boundCheck = boundCheck > bound * 1E-10 ? boundCheck : bound * 1E-10;
                boundCheck = 1.0 / exp(boundCheck * boundCheck * boundCheck);
                boundCheck *= copysign(boundCheck, -x);
                rn += boundCheck;
            }
            s->navErrors[i] = x + NAV_P[i] * rn;
        }
        s->navPrevTime = now;
    }
    v3Add(s->scState.r, &s->navErrors[0], s->navTrans.r);
    v3Add(s->scState.v, &s->navErrors[3], s->navTrans.v);
    orc_addMRP(s->scState.sigma, &s->navErrors[6], s->navAtt.sigma);
    v3Add(s->scState.omega, &s->navErrors[9], s->navAtt.omega);
    double sc2sun[3], BN[3][3], sunTrue_B[3], OT[3][3];
    v3Subtract(s->sunMsg.pos, s->scState.r, sc2sun);
    v3Normalize(sc2sun, sc2sun);
    orc_MRP2C(s->scState.sigma, BN);
    m33MultV3(BN, sc2sun, sunTrue_B);
    orc_MRP2C(&s->navErrors[12], OT);
    m33MultV3(OT, sunTrue_B, s->navAtt.sun_B);
    s->navAtt.h.written = 1; s->navTrans.h.written = 1;
}
static void camera_update(orc_opnav_sim *s, uint64_t now)
{ /* CameraTask, 60 s (OND:62, :132-133): the frame the renderer would be asked for -- truth state at `now` */
    if (!s->cameraIsOn) return;
    s->image.valid = 1; s->image.timeTag = now;
    v3Copy(s->scState.r, s->image.r); v3Copy(s->scState.sigma, s->image.sigma);
    s->image.h.written = 1; s->n_images++;
}

/* ---- FSW models ---- */
static void hillPoint_update(orc_opnav_sim *s)
{ /* [BSK: hillPoint.c]; planet ephemeris = Mars at the origin of the zero base (ONF:288) */
    double dcm_RN[3][3], h[3];
    const double *relPos = s->navTrans.r, *relVel = s->navTrans.v;
    v3Normalize(relPos, dcm_RN[0]);
    v3Cross(relPos, relVel, h);
    v3Normalize(h, dcm_RN[2]);
    v3Cross(dcm_RN[2], dcm_RN[0], dcm_RN[1]);
    orc_C2MRP(dcm_RN, s->attRef.sigma_RN);
    double rm = v3Norm(relPos), hm = v3Norm(h), dfdt, ddfdt2;
    if (rm > 1.) { dfdt = hm / (rm * rm); ddfdt2 = -2.0 * v3Dot(relVel, dcm_RN[0]) / rm * dfdt; }
    else { dfdt = 0.; ddfdt2 = 0.; }
    double omega_RN_R[3] = {0, 0, dfdt}, domega_RN_R[3] = {0, 0, ddfdt2};
    m33tMultV3(dcm_RN, omega_RN_R, s->attRef.omega_RN_N);
    m33tMultV3(dcm_RN, domega_RN_R, s->attRef.domega_RN_N);
    s->attRef.h.written = 1;
}
static void trackingErrorCam_update(orc_opnav_sim *s)
{ /* [BSK: attTrackingError.c] with sigma_R0R = C2MRP(M3 M2 M_cam) (ONF:345-356) */
    double sigma_RR0[3], sigma_RN[3], BN[3][3];
    v3Scale(-1.0, s->sigma_R0R, sigma_RR0);
    orc_addMRP(s->attRef.sigma_RN, sigma_RR0, sigma_RN);
    orc_subMRP(s->navAtt.sigma, sigma_RN, s->attGuid.sigma_BR);
    orc_MRP2C(s->navAtt.sigma, BN);
    m33MultV3(BN, s->attRef.omega_RN_N, s->attGuid.omega_RN_B);
    v3Subtract(s->navAtt.omega, s->attGuid.omega_RN_B, s->attGuid.omega_BR_B);
    m33MultV3(BN, s->attRef.domega_RN_N, s->attGuid.domega_RN_B);
    s->attGuid.h.written = 1;
}
static void sunSafePoint_update(orc_opnav_sim *s)
{ /* [BSK: sunSafePoint.c] [M]: sHatBdyCmd = (0,0,1) (ONF:295), minUnitMag = smallAngle = sunAxisSpinRate = 0 */
    double sHat[3] = {0, 0, 0}, omega_BN_B[3] = {0, 0, 0};
    if (s->sunPoint.h.written) v3Copy(s->sunPoint.sun_B, sHat);
    if (s->navAtt.h.written) v3Copy(s->navAtt.omega, omega_BN_B);
    double sNorm = v3Norm(sHat);
    if (sNorm > 0.0) {
        double ct = v3Dot(s->sHatBdyCmd, sHat) / sNorm;
        ct = fabs(ct) > 1.0 ? ct / fabs(ct) : ct;
        double err = acos(ct), e_hat[3];
        if (err < 0.0) v3SetZero(s->attGuid.sigma_BR);
        else {
            if (PI_D - err < 0.0) v3Copy(s->eHat180_B, e_hat);
            else v3Cross(sHat, s->sHatBdyCmd, e_hat);
            v3Normalize(e_hat, e_hat);
            v3Scale(tan(err * 0.25), e_hat, s->attGuid.sigma_BR);
            double m = v3Dot(s->attGuid.sigma_BR, s->attGuid.sigma_BR);
            if (m > 1.0) v3Scale(-1. / m, s->attGuid.sigma_BR, s->attGuid.sigma_BR);
        }
        v3SetZero(s->attGuid.omega_RN_B);           /* sunAxisSpinRate = 0 */
    } else {
        v3SetZero(s->attGuid.sigma_BR);
        v3SetZero(s->attGuid.omega_RN_B);           /* search rate omega_RN_B = 0 */
    }
    v3Subtract(omega_BN_B, s->attGuid.omega_RN_B, s->attGuid.omega_BR_B);
    v3SetZero(s->attGuid.domega_RN_B);
    s->attGuid.h.written = 1;
}
static void cssWlsEst_update(orc_opnav_sim *s)
{ /* [BSK: cssWlsEst.c] [M]: sensorUseThresh = 0, useWeights = 0, CBias = 1 (ONF:373); three or more sensors ->
     least squares d = (H^T H)^-1 H^T y, one or two -> minimum norm d = H^T (H H^T)^-1 y, none -> zero heading */
    double H[NCSS][3], y[NCSS]; int n = 0;
    for (int i = 0; i < NCSS; i++)
        if (s->cssMsg.h.written && s->cssMsg.y[i] > 0.0) { v3Copy(s->cssN[i], H[n]); y[n] = s->cssMsg.y[i]; n++; }
    double d[3] = {0, 0, 0};
    if (n >= 3) {
        double HtH[3][3] = {{0}}, Hty[3] = {0, 0, 0}, inv[3][3];
        for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) { Hty[a] += H[i][a] * y[i]; for (int b = 0; b < 3; b++) HtH[a][b] += H[i][a] * H[i][b]; }
        m33Inverse(HtH, inv);
        m33MultV3(inv, Hty, d);
    } else if (n == 2) {
        double a = v3Dot(H[0], H[0]), b = v3Dot(H[0], H[1]), c = v3Dot(H[1], H[1]), det = a * c - b * b;
        double l0 = (c * y[0] - b * y[1]) / det, l1 = (a * y[1] - b * y[0]) / det;
        for (int k = 0; k < 3; k++) d[k] = H[0][k] * l0 + H[1][k] * l1;
    } else if (n == 1) {
        v3Scale(y[0] / v3Dot(H[0], H[0]), H[0], d);
    }
    if (n > 0) v3Normalize(d, s->sunPoint.sun_B); else v3SetZero(s->sunPoint.sun_B);
    s->sunPoint.h.written = 1;
}
static void mrpFeedbackRWs_update(orc_opnav_sim *s)
{ /* [BSK: MRP_Feedback.c] with the wheel momentum feed-forward (ONF:399-409): K 3.5, P 30, Ki -1 (integral off) */
    AttGuidMsg g; memset(&g, 0, sizeof(g));
    if (s->attGuid.h.written) g = s->attGuid;
    double omega_BN_B[3], Lr[3], v1[3], v4[3], v7[3], v8[3], v9[3], v10[3];
    v3Add(g.omega_BR_B, g.omega_RN_B, omega_BN_B);
    v3Scale(s->K, g.sigma_BR, Lr);
    v3Scale(s->P, g.omega_BR_B, v1);
    v3Add(v1, Lr, Lr);
    m33MultV3(s->ISC, omega_BN_B, v4);
    for (int i = 0; i < NRW; i++) {
        double Om = s->rwSpeeds.h.written ? s->rwSpeeds.Omega[i] : 0.0;
        double hs = s->Js * (v3Dot(omega_BN_B, s->gs[i]) + Om), t[3];
        v3Scale(hs, s->gs[i], t);
        v3Add(t, v4, v4);
    }
    v3Cross(g.omega_RN_B, v4, v7);
    v3Subtract(Lr, v7, Lr);
    v3Cross(omega_BN_B, g.omega_RN_B, v8);
    v3Subtract(g.domega_RN_B, v8, v9);
    m33MultV3(s->ISC, v9, v10);
    v3Subtract(Lr, v10, Lr);
    v3Scale(-1.0, Lr, s->cmdTorque.Lr);
    s->cmdTorque.h.written = 1;
}
static void rwMotorTorque_update(orc_opnav_sim *s)
{ /* [BSK: rwMotorTorque.c]: us = CGs^T (CGs CGs^T)^-1 (-Lr), controlAxes_B = I (ONF:437-446) */
    double Lr_B[3] = {0, 0, 0}, M[3][3] = {{0}}, t[3];
    if (s->cmdTorque.h.written) v3Copy(s->cmdTorque.Lr, Lr_B);
    v3Scale(-1.0, Lr_B, Lr_B);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < NRW; k++) M[i][j] += s->gs[k][i] * s->gs[k][j];
    m33Inverse(M, M);
    m33MultV3(M, Lr_B, t);
    for (int i = 0; i < NRW; i++) s->rwCmd.u[i] = v3Dot(s->gs[i], t);
    s->rwCmd.h.written = 1;
}
static void imageProcessing_update(orc_opnav_sim *s, uint64_t now)
{ /* stand-in for houghCircles (ONF:452-468): a NEW valid frame yields one circle (+ optional pixel noise) */
    s->circles.valid = 0; s->circles.h.written = 1;
    if (!(s->image.h.written && s->image.valid && s->image.timeTag >= now)) return;
    double BN[3][3], mr[3], r_C[3], c[3];
    orc_MRP2C(s->image.sigma, BN);
    v3Scale(-1.0, s->image.r, mr);
    m33MultV3(BN, mr, r_C);                      /* sigma_CB = 0 (OND:134-135): camera frame == body frame */
    if (!orc_opnav_project_circle(r_C, REQ_MARS_KM * 1000., c)) return;
    if (s->cfg.pixel_noise_std > 0.0) {
        double n[4];
        orc_opnav_normals(s->cfg.seed, s->env_index, s->episode, (uint32_t)(now / s->dyn_ns), 2, 0, n);
        for (int k = 0; k < 3; k++) c[k] += s->cfg.pixel_noise_std * n[k];
    }
    s->circles.valid = 1; s->circles.timeTag = s->image.timeTag; s->circles.unc = s->cfg.circle_unc;
    v3Copy(c, s->circles.c);
}
static void pixelLine_update(orc_opnav_sim *s)
{
    s->opnav.h.written = 1;
    if (!s->circles.valid) { s->opnav.valid = 0; return; }
    double CN[3][3];
    orc_MRP2C(s->navAtt.sigma, CN);              /* attInMsgName = simple_att_nav_output (ONF:473), dcm_CB = I */
    orc_opnav_pixel_line(s->circles.c, s->circles.unc, CN, s->opnav.r_BN_N, s->opnav.covar_N);
    s->opnav.valid = 1; s->opnav.timeTag = s->circles.timeTag;
}
static void relativeOD_update(orc_opnav_sim *s, uint64_t now)
{ /* Update_relODuKF [M] */
    double newTimeTag = (double)s->opnav.timeTag * NANO2SEC;
    if (s->opnav.h.written && s->opnav.valid && newTimeTag >= s->ukf.timeTag) {
        orc_ukf_time_update(&s->ukf, newTimeTag);
        orc_ukf_meas_update(&s->ukf, s->opnav.r_BN_N, s->opnav.covar_N);
        s->n_meas++;
        s->opnav.valid = 0;                      /* a measurement is consumed once (its timeTag can no longer be >= filter time
                                                    except at dt = 0, which this guard removes) */
    }
    newTimeTag = now * NANO2SEC;
    if (newTimeTag > s->ukf.timeTag) orc_ukf_time_update(&s->ukf, newTimeTag);
}

static void single_step(orc_opnav_sim *s, uint64_t now)
{
    /* DynamicsProcess (priority 100): DynamicsTask (prio 1000), then CameraTask (prio 999, 60 s) */
    rw_update(s);
    css_update(s);
    eclipse_update(s);
    sc_update(s, now);
    spice_update(s, now);
    simple_nav_update(s, now);
    if (now % (60ull * 1000000000ull) == 0) camera_update(s, now);
    /* FSWProcess (priority 10): tasks by priority 20, 20, 15, 5 (ONF:108-122) */
    if (s->t_opNavPointCheat) { hillPoint_update(s); trackingErrorCam_update(s); }
    if (s->t_sunSafe) { sunSafePoint_update(s); cssWlsEst_update(s); relativeOD_update(s, now); }
    if (s->t_mrpRW) { mrpFeedbackRWs_update(s); rwMotorTorque_update(s); }
    if (s->t_opNavOD) { imageProcessing_update(s, now); pixelLine_update(s); relativeOD_update(s, now); }
}

void orc_opnav_default_cfg(orc_opnav_cfg *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->dynRate = 1.0; cfg->fswRate = 1.0; cfg->step_duration_min = 50.0;
    cfg->nav_noise = 1; cfg->camera_reenable = 0; cfg->pixel_noise_std = 0.5; cfg->circle_unc = 0.25;
    cfg->seed = 0; cfg->numModes = 50;
}
void orc_opnav_reference_orbit(orc_opnav_ic *ic)
{
    memset(ic, 0, sizeof(*ic));
    orc_elem2rv(MU_MARS_DYN, 18000 * 1E3, 0.6, 10 * D2R, 25. * D2R, 190. * D2R, 80. * D2R, ic->rN, ic->vN);   /* ONS:173-181 */
}
orc_opnav_sim *orc_opnav_create(const orc_opnav_ic *ic, const orc_opnav_cfg *cfg, uint64_t env_index, uint64_t episode)
{
    orc_opnav_sim *s = (orc_opnav_sim *)calloc(1, sizeof(*s));
    s->cfg = *cfg; s->ic = *ic; s->env_index = env_index; s->episode = episode;
    s->dyn_ns = (uint64_t)(cfg->dynRate * 1e9 + 0.5);
    s->mHub = 750.0;                                                       /* OND:180 */
    s->IHub[0][0] = 900.; s->IHub[1][1] = 800.; s->IHub[2][2] = 600.;      /* OND:177-179 */
    memcpy(s->ISC, s->IHub, sizeof(s->ISC));                               /* ONF:414 */
    v3Copy(ic->rN, s->r); v3Copy(ic->vN, s->v);                            /* sigma = omega = 0 (ONS:189-194) */
    for (int i = 0; i < NRW; i++) {                                        /* OND:269-293, ONF:418-432 */
        double el = 40.0 * D2R, az = (45.0 + 90.0 * i) * D2R;
        s->gs[i][0] = cos(el) * cos(az); s->gs[i][1] = cos(el) * sin(az); s->gs[i][2] = sin(el);
    }
    s->Omega_max = 6000.0 * RPM; s->u_max = 0.2; s->Js = 50. / s->Omega_max;
    { const double n[NCSS][3] = {{0.0, 0.707107, 0.707107}, {0.707107, 0., 0.707107}, {0.0, -0.707107, 0.707107},
                                 {-0.707107, 0., 0.707107}, {0.0, -0.965926, -0.258819}, {-0.707107, -0.353553, -0.612372},
                                 {0., 0.258819, -0.965926}, {0.707107, -0.353553, -0.612372}};   /* OND:341-350 */
      memcpy(s->cssN, n, sizeof(n)); }
    s->sHatBdyCmd[2] = 1.0; s->eHat180_B[0] = 1.0;
    s->K = 3.5; s->P = 30.0; s->Ki = -1;
    { /* sigma_R0R = C2MRP(euler1(90) euler2(90) MRP2C(cameraMRP_CB = 0)) (ONF:350-355) */
      double M[3][3] = {{0, 0, -1}, {1, 0, 0}, {0, -1, 0}};
      orc_C2MRP(M, s->sigma_R0R); }
    { /* relativeODuKF (ONF:495-527; stateInit/qNoise/noiseSF overridden at ONS:189, :196-200) */
      double st[6], P0[36], Q[36];
      memset(P0, 0, sizeof(P0)); memset(Q, 0, sizeof(Q));
      for (int k = 0; k < 3; k++) {
          st[k] = ic->rN[k] + ic->rError[k]; st[3 + k] = ic->vN[k] + ic->vError[k];
          P0[k * 7] = 1. * 1E6; P0[(k + 3) * 7] = 0.02 * 1E6;
          Q[k * 7] = 1E-3 * 1E-3; Q[(k + 3) * 7] = 1E-4 * 1E-4;
      }
      orc_ukf_init(&s->ukf, st, P0, Q, MU_MARS_FSW, 5.0); }
    s->cameraIsOn = 1;                                                     /* OND:130 */
    s->first_run = 1; s->modeCounter = 0; s->simTime = 0.0;
    spice_update(s, 0);                                                    /* SpiceInterface Reset publishes the planets */
    s->next_ns = 0;
    return s;
}
void orc_opnav_destroy(orc_opnav_sim *s) { free(s); }

int orc_opnav_run_sim(orc_opnav_sim *s, int action, double obs[4], double debug[12])
{ /* ONS:225-299 */
    s->modeCounter += 1;
    if (action == 0) {
        s->t_opNavPointCheat = 1; s->t_mrpRW = 1; s->t_opNavOD = 1; s->t_sunSafe = 0;
        if (s->cfg.camera_reenable) s->cameraIsOn = 1;     /* ONS:239 is commented out in the reference */
    } else if (action == 1) {
        s->cameraIsOn = 0;                                  /* ONS:249 */
        s->t_opNavPointCheat = 0; s->t_opNavOD = 0; s->t_sunSafe = 1; s->t_mrpRW = 1;
    }
    if (s->first_run) {
        /* modeRequest = 'OpNavOD' (ONS:157) is still pending: the 'OpNavOD' event (ONF:219-224) fires at the first
           ExecuteSimulation, AFTER the task switching above, and selects the OpNavOD task set whatever the action */
        s->t_opNavPointCheat = 1; s->t_mrpRW = 1; s->t_opNavOD = 1; s->t_sunSafe = 0;
        s->first_run = 0;
    }
    s->simTime += s->cfg.step_duration_min;
    uint64_t stop = (uint64_t)(s->simTime * 60.0 * 1e9 + 0.5);        /* mc.min2nano */
    while (s->next_ns <= stop) { single_step(s, s->next_ns); s->next_ns += s->dyn_ns; }   /* stop time inclusive */
    /* observation (ONS:263-293), last logged record = the final tick */
    const double *navState = s->ukf.state, *navCovar = s->ukf.covar;
    double covarVec[3] = {sqrt(navCovar[0]), sqrt(navCovar[1 + NST]), sqrt(navCovar[2 + 2 * NST])};
    double BN[3][3], rhat[3], pos_B[3], sunHeadNorm[3];
    orc_MRP2C(s->scState.sigma, BN);
    double nr = v3Norm(navState);
    v3Scale(1.0 / nr, navState, rhat);
    m33MultV3(BN, rhat, pos_B);
    v3Scale(-1.0, pos_B, pos_B);
    v3Scale(1.0 / v3Norm(s->navAtt.sun_B), s->navAtt.sun_B, sunHeadNorm);
    s->obs[0] = v3Dot(pos_B, sunHeadNorm);
    for (int k = 0; k < 3; k++) s->obs[1 + k] = covarVec[k] / nr;
    for (int k = 0; k < 3; k++) { s->debug[k] = navState[k]; s->debug[3 + k] = s->scState.r[k]; s->debug[6 + k] = s->scState.v[k]; s->debug[9 + k] = s->scState.sigma[k]; }
    memcpy(obs, s->obs, sizeof(s->obs));
    if (debug) memcpy(debug, s->debug, sizeof(s->debug));
    return s->modeCounter >= s->cfg.numModes;
}
void orc_opnav_get_state(const orc_opnav_sim *s, orc_opnav_state *o)
{
    memset(o, 0, sizeof(*o));
    v3Copy(s->r, o->r_BN_N); v3Copy(s->v, o->v_BN_N); v3Copy(s->sigma, o->sigma_BN); v3Copy(s->omega, o->omega_BN_B);
    memcpy(o->Omega, s->Omega, sizeof(o->Omega)); memcpy(o->u_current, s->u_current, sizeof(o->u_current));
    memcpy(o->navErrors, s->navErrors, sizeof(o->navErrors));
    v3Copy(s->navTrans.r, o->nav_r); v3Copy(s->navTrans.v, o->nav_v); v3Copy(s->navAtt.sigma, o->nav_sigma);
    v3Copy(s->navAtt.omega, o->nav_omega); v3Copy(s->navAtt.sun_B, o->nav_sun_B);
    v3Copy(s->attGuid.sigma_BR, o->sigma_BR); v3Copy(s->attGuid.omega_BR_B, o->omega_BR_B); v3Copy(s->cmdTorque.Lr, o->Lr);
    memcpy(o->rwCmd, s->rwCmd.u, sizeof(o->rwCmd)); memcpy(o->css, s->cssMsg.y, sizeof(o->css));
    v3Copy(s->sunPoint.sun_B, o->sun_point); o->shadow = s->eclipseMsg.shadow;
    memcpy(o->filt_state, s->ukf.state, sizeof(o->filt_state)); memcpy(o->filt_covar, s->ukf.covar, sizeof(o->filt_covar));
    memcpy(o->filt_sBar, s->ukf.sBar, sizeof(o->filt_sBar)); o->filt_time = s->ukf.timeTag;
    v3Copy(s->opnav.r_BN_N, o->meas_r); memcpy(o->meas_covar, s->opnav.covar_N, sizeof(o->meas_covar));
    v3Copy(s->circles.c, o->circle);
    o->n_meas = s->n_meas; o->n_bad = s->ukf.n_bad; o->n_images = s->n_images; o->mrp_switch_count = s->MRPSwitchCount;
    o->camera_on = s->cameraIsOn; o->mode = s->t_sunSafe ? 1 : 0; o->modeCounter = s->modeCounter;
    o->sim_nanos = s->next_ns - s->dyn_ns;
}

/* ================================ gym layer (ONE:55-125) ======================================== */
struct orc_opnav_env {
    orc_opnav_cfg cfg; orc_opnav_sim *sim;
    int curr_step, episode_over, max_length; double reward_total, reward_mult;
    double obs[4], debug[12];
};
orc_opnav_env *orc_opnav_env_create(const orc_opnav_cfg *cfg)
{
    orc_opnav_env *e = (orc_opnav_env *)calloc(1, sizeof(*e));
    e->cfg = *cfg; e->max_length = 40; e->reward_mult = 1.0;      /* ONE:23, :32 */
    return e;
}
void orc_opnav_env_destroy(orc_opnav_env *e) { if (e->sim) orc_opnav_destroy(e->sim); free(e); }
orc_opnav_sim *orc_opnav_env_sim(orc_opnav_env *e) { return e->sim; }
void orc_opnav_env_set_max_length(orc_opnav_env *e, int max_length) { e->max_length = max_length; }
/* opNavEnvironment.py:106-109: info['episode'] = {'r': reward_total, 'l': curr_step}, assembled before curr_step += 1 (:122) */
void orc_opnav_env_episode(const orc_opnav_env *e, double *r, int *l) { *r = e->reward_total; *l = e->curr_step - 1; }
void orc_opnav_env_reset(orc_opnav_env *e, const orc_opnav_ic *ic, uint64_t env_index, uint64_t episode, double ob[4])
{ /* ONE:153-168: a fresh simulator; the initial observation is zeros (ONS:152) */
    if (e->sim) orc_opnav_destroy(e->sim);
    e->sim = orc_opnav_create(ic, &e->cfg, env_index, episode);
    e->episode_over = 0; e->curr_step = 0; e->reward_total = 0;
    memset(e->obs, 0, sizeof(e->obs)); memset(e->debug, 0, sizeof(e->debug));
    if (ob) memset(ob, 0, 4 * sizeof(double));
}
void orc_opnav_env_step(orc_opnav_env *e, int action, orc_opnav_out *out)
{
    out->reason = 0;
    if (e->curr_step >= e->max_length) { e->episode_over = 1; out->reason |= 1; }     /* ONE:94-95 */
    int sim_over = orc_opnav_run_sim(e->sim, action, e->obs, e->debug);              /* ONE:98 */
    double reward = 0, nav[3], real[3] = {e->debug[3], e->debug[4], e->debug[5]};     /* ONE:139-152 */
    for (int k = 0; k < 3; k++) nav[k] = e->debug[k] - real[k];
    double nr = 1. / v3Norm(real);
    for (int k = 0; k < 3; k++) nav[k] *= nr;
    if (action == 1) reward = fabs(e->reward_mult / (1. + pow(v3Norm(nav), 2.0)));
    e->reward_total += reward;
    if (sim_over) { e->episode_over = 1; out->reason |= 2; }                          /* ONE:104-106 */
    memcpy(out->ob, e->obs, sizeof(e->obs)); memcpy(out->debug, e->debug, sizeof(e->debug));
    out->reward = reward; out->done = e->episode_over;
    e->curr_step += 1;
}
void orc_opnav_env_step_batch(orc_opnav_env **envs, int n, const int *actions, orc_opnav_out *outs, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (int i = 0; i < n; i++) orc_opnav_env_step(envs[i], actions[i], &outs[i]);
    (void)nthreads;
}
