/*
 * mpm_oracle.c -- CPU restatement of MPMAvatar's per-substep MPM solver.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mpmavatar_b200/ may link, import
 * or execute this file; it exists so tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can check and time the
 * reference algorithm on the host.
 *
 * PINNING: the reference (KAISTChangmin/MPMAvatar) ships no tests, golden
 * vectors or fixtures for this path and its runtime,
 * warp_lang==0.10.1 (requirements.txt:36), is a third-party dependency that
 * is neither vendored under /root/reference nor installable offline.  This
 * file is a line-by-line restatement of the reference's Python/Warp kernels
 * (each function cites the file:line it follows) and is pinned by
 *   (a) the analytic known-answer tests in tests/test_oracle_kat.py, and
 *   (b) golden vectors produced by executing the reference's OWN source
 *       (warp_mpm/mpm_solver.py, mpm_utils.py, mpm_data_structure.py,
 *       unmodified) under oracle/warp_emu.py, a sequential Python stand-in
 *       for the Warp API: tests/golden/make_golden.py -> the .npz files beside it,
 *       checked by tests/test_golden.py (fp64 build: 1e-9 relative on every
 *       state and grid field, all materials, collider, mover, plane).
 * What remains an assumption is the convention of the two third-party
 * numerical routines below (their source is not readable offline).
 *
 * Third-party pieces restated from their published behaviour (warp-lang 0.10.x):
 *   wp.qr3   -> Givens (rotation) QR, det Q = +1; only its sign-normalised
 *               result is consumed (mpm_utils.py:112-123), which equals
 *               Gram-Schmidt with q3 = q1 x q2.
 *   wp.svd3  -> A = U diag(s) V^T with det U = det V = +1, s sorted
 *               descending, the sign of det A carried by s[2]
 *               (McAdams et al. convention).  Restated with a Jacobi
 *               eigen-solve of A^T A carried out in double precision.
 *   wp.mat33(vec,vec,vec) -> the three vectors are COLUMNS.
 *   wp.int() -> truncation toward zero.   wp.normalize(0) -> 0.
 *   wp.mesh_eval_face_normal -> normalize(cross(q-p, r-p)).
 *
 * Layout: dense n_grid^3 grid exactly like the reference; particle order
 * [elements | traditional | vertices] (train_material_params.py:387).
 * Build: see oracle/Makefile (fp32 default; -DORACLE_FP64 for the drift
 * envelope; -fopenmp for the multi-core CPU baseline).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_FP64
typedef double real;
#define RSQRT(x) sqrt(x)
#define RLOG(x) log(x)
#define REXP(x) exp(x)
#define RABS(x) fabs(x)
#else
typedef float real;
#define RSQRT(x) sqrtf(x)
#define RLOG(x) logf(x)
#define REXP(x) expf(x)
#define RABS(x) fabsf(x)
#endif

#define R(x) ((real)(x))

/* boundary-condition record, one per grid_postprocess entry (mpm_solver.py:564-658,
 * 929-984, 986-1053, 1330-1355) */
enum { BC_SURFACE = 0, BC_CUBOID = 1, BC_BBOX = 2, BC_MASK = 3 };
typedef struct {
    int kind;
    int surface_type; /* 0 sticky, 1 slip, 11 cut, 2 other */
    int reset;
    int pad_;
    real point[3];
    real normal[3];
    real size[3];
    real velocity[3];
    real friction;
    real start_time;
    real end_time;
    const int *mask; /* BC_MASK: n^3 ints */
} OrcBC;

typedef struct {
    /* sizes */
    int n_particles, n_elements, n_vertices, n_grid;
    /* model scalars (mpm_data_structure.py:610-645, 686-715) */
    real grid_lim, dx, inv_dx;
    int material;
    int hardening_i; /* model.hardening == 1 test (mpm_utils.py:249) */
    real friction_coeff, alpha;
    real g[3];
    real rpic_damping, grid_v_damping_scale;
    real xi, plastic_viscosity, softening;
    /* particle arrays */
    real *x, *v, *C;             /* N*3, N*3, N*9 */
    real *F, *F_trial, *stress;  /* Nnv*9 */
    real *d;                     /* Ne*9 */
    real *R_inv;                 /* Ne*3 */
    real *faces;                 /* Ne*3, vertex-local indices stored as reals */
    real *vertex_force;          /* Nv*3 */
    real *vol, *mass;            /* N */
    int *traditional, *vertices, *elements, *selection; /* N */
    real *mu, *lam, *gamma, *kappa, *yield_stress;      /* N */
    /* dense grid */
    real *grid_m, *grid_v_in, *grid_v_out; /* n^3, n^3*3, n^3*3 */
    /* mesh collider (mpm_solver.py:805-919); has_collider=0 -> skipped */
    int has_collider;
    int n_mesh_v, n_mesh_f;
    real collider_friction;
    const int *mesh_faces;        /* Mf*3 */
    real *mesh_points, *mesh_velocities; /* Mv*3 (wp.Mesh points / velocities) */
    real *col_weight, *col_v_in, *col_v_out, *col_normal; /* n^3 (x3) */
    /* particle mover (mpm_solver.py:661-802) */
    int has_mover;
    int num_joint_v, num_joint_f;
    real *mov_weight, *mov_velocity;
    /* ordered BC list */
    int n_bc;
    OrcBC *bc;
} OrcSim;

/* ---------------------------------------------------------------- mat helpers
 * 3x3 matrices are row-major real[9], m[3*r+c], like wp.mat33 scalars ctor. */
static inline void mat_mul(const real *a, const real *b, real *o) {
    real t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    memcpy(o, t, sizeof t);
}
static inline void mat_T(const real *a, real *o) {
    real t[9] = {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]};
    memcpy(o, t, sizeof t);
}
static inline real mat_det(const real *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
           m[2] * (m[3] * m[7] - m[4] * m[6]);
}
static inline real vlen(const real *a) { return RSQRT(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
/* wp.normalize: returns 0 for the zero vector */
static inline void vnormalize(const real *a, real *o) {
    real l = vlen(a);
    if (l > R(0)) { o[0] = a[0] / l; o[1] = a[1] / l; o[2] = a[2] / l; }
    else { o[0] = o[1] = o[2] = R(0); }
}

/* ------------------------------------------------------------------ wp.qr3
 * Rotation QR of a 3x3 (columns a1,a2,a3), then the reference's sign
 * normalisation (mpm_utils.py:109-123 / 181-195).  The result has
 * R00>=0, R11>=0 and det Q=+1, i.e. Gram-Schmidt with q3=q1xq2. */
void orc_qr3_signed(const real *A, real *Q, real *Rm) {
    real a1[3] = {A[0], A[3], A[6]}, a2[3] = {A[1], A[4], A[7]}, a3[3] = {A[2], A[5], A[8]};
    real q1[3], q2[3], q3[3], u2[3];
    real r00 = vlen(a1);
    for (int i = 0; i < 3; i++) q1[i] = a1[i] / r00;
    real r01 = q1[0] * a2[0] + q1[1] * a2[1] + q1[2] * a2[2];
    for (int i = 0; i < 3; i++) u2[i] = a2[i] - r01 * q1[i];
    real r11 = vlen(u2);
    for (int i = 0; i < 3; i++) q2[i] = u2[i] / r11;
    q3[0] = q1[1] * q2[2] - q1[2] * q2[1];
    q3[1] = q1[2] * q2[0] - q1[0] * q2[2];
    q3[2] = q1[0] * q2[1] - q1[1] * q2[0];
    real r02 = q1[0] * a3[0] + q1[1] * a3[1] + q1[2] * a3[2];
    real r12 = q2[0] * a3[0] + q2[1] * a3[1] + q2[2] * a3[2];
    real r22 = q3[0] * a3[0] + q3[1] * a3[1] + q3[2] * a3[2];
    for (int i = 0; i < 3; i++) { Q[3 * i] = q1[i]; Q[3 * i + 1] = q2[i]; Q[3 * i + 2] = q3[i]; }
    Rm[0] = r00; Rm[1] = r01; Rm[2] = r02;
    Rm[3] = 0;   Rm[4] = r11; Rm[5] = r12;
    Rm[6] = 0;   Rm[7] = 0;   Rm[8] = r22;
}

/* ------------------------------------------------------------------ wp.svd3
 * A = U diag(s) V^T, det U = det V = +1, s[0]>=s[1]>=|s[2]|. */
void orc_svd3(const real *Ain, real *Uo, real *So, real *Vo) {
    double A[9], B[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) A[i] = (double)Ain[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            B[3 * i + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = B[1] * B[1] + B[2] * B[2] + B[5] * B[5];
        double dia = B[0] * B[0] + B[4] * B[4] + B[8] * B[8];
        if (off <= 1e-34 * dia || off == 0.0) break;
        static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
        for (int k = 0; k < 3; k++) {
            int p = PQ[k][0], q = PQ[k][1];
            double apq = B[3 * p + q];
            if (apq == 0.0) continue;
            double theta = (B[3 * q + q] - B[3 * p + p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int r = 0; r < 3; r++) { /* B <- B J */
                double bp = B[3 * r + p], bq = B[3 * r + q];
                B[3 * r + p] = c * bp - s * bq;
                B[3 * r + q] = s * bp + c * bq;
            }
            for (int r = 0; r < 3; r++) { /* B <- J^T B */
                double bp = B[3 * p + r], bq = B[3 * q + r];
                B[3 * p + r] = c * bp - s * bq;
                B[3 * q + r] = s * bp + c * bq;
            }
            for (int r = 0; r < 3; r++) {
                double vp = V[3 * r + p], vq = V[3 * r + q];
                V[3 * r + p] = c * vp - s * vq;
                V[3 * r + q] = s * vp + c * vq;
            }
        }
    }
    double lam[3] = {B[0], B[4], B[8]};
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (lam[idx[j]] < lam[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
    double Vs[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) Vs[3 * r + c] = V[3 * r + idx[c]];
    double detV = Vs[0] * (Vs[4] * Vs[8] - Vs[5] * Vs[7]) - Vs[1] * (Vs[3] * Vs[8] - Vs[5] * Vs[6]) +
                  Vs[2] * (Vs[3] * Vs[7] - Vs[4] * Vs[6]);
    if (detV < 0) for (int r = 0; r < 3; r++) Vs[3 * r + 2] = -Vs[3 * r + 2];
    double AV[3][3]; /* AV[c] = A * v_c */
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            AV[c][r] = A[3 * r] * Vs[c] + A[3 * r + 1] * Vs[3 + c] + A[3 * r + 2] * Vs[6 + c];
    double u0[3], u1[3], u2[3];
    double s0 = sqrt(AV[0][0] * AV[0][0] + AV[0][1] * AV[0][1] + AV[0][2] * AV[0][2]);
    if (s0 > 1e-300) { for (int r = 0; r < 3; r++) u0[r] = AV[0][r] / s0; }
    else { u0[0] = 1; u0[1] = 0; u0[2] = 0; }
    double dp = u0[0] * AV[1][0] + u0[1] * AV[1][1] + u0[2] * AV[1][2];
    double w1[3] = {AV[1][0] - dp * u0[0], AV[1][1] - dp * u0[1], AV[1][2] - dp * u0[2]};
    double n1 = sqrt(w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2]);
    if (n1 > 1e-14 * (s0 > 0 ? s0 : 1.0)) { for (int r = 0; r < 3; r++) u1[r] = w1[r] / n1; }
    else { /* any unit vector orthogonal to u0 */
        int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2)
                                          : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[m] = 1;
        double dd = u0[m];
        double w[3] = {e[0] - dd * u0[0], e[1] - dd * u0[1], e[2] - dd * u0[2]};
        double nn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        for (int r = 0; r < 3; r++) u1[r] = w[r] / nn;
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    double s1 = u1[0] * AV[1][0] + u1[1] * AV[1][1] + u1[2] * AV[1][2];
    double s2 = u2[0] * AV[2][0] + u2[1] * AV[2][1] + u2[2] * AV[2][2];
    for (int r = 0; r < 3; r++) {
        Uo[3 * r] = (real)u0[r]; Uo[3 * r + 1] = (real)u1[r]; Uo[3 * r + 2] = (real)u2[r];
    }
    for (int i = 0; i < 9; i++) Vo[i] = (real)Vs[i];
    So[0] = (real)s0; So[1] = (real)s1; So[2] = (real)s2;
}

static inline void diag_sandwich(const real *U, const real *dg, const real *V, real *o) {
    /* U * diag(dg) * V^T */
    real UD[9], VT[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) UD[3 * r + c] = U[3 * r + c] * dg[c];
    mat_T(V, VT);
    mat_mul(UD, VT, o);
}

/* --------------------------------------------------- constitutive functions */
/* mpm_utils.py:8-15 */
static void kirchoff_stress_FCR(const real *F, const real *U, const real *V, real J, real mu,
                                real lam, real *out) {
    real VT[9], Rm[9], FT[9], D[9], t[9];
    mat_T(V, VT); mat_mul(U, VT, Rm); mat_T(F, FT);
    for (int i = 0; i < 9; i++) D[i] = F[i] - Rm[i];
    mat_mul(D, FT, t);
    for (int i = 0; i < 9; i++) out[i] = R(2.0) * mu * t[i];
    real p = lam * J * (J - R(1.0));
    out[0] += p; out[4] += p; out[8] += p;
}
/* mpm_utils.py:50-66 */
static void kirchoff_stress_StVK(const real *F, const real *U, const real *V, const real *sig_in,
                                 real mu, real lam, real *out) {
    real sig[3], eps[3], tau[3], t[9], FT[9];
    for (int i = 0; i < 3; i++) sig[i] = sig_in[i] > R(0.01) ? sig_in[i] : R(0.01);
    for (int i = 0; i < 3; i++) eps[i] = RLOG(sig[i]);
    real sum = RLOG(sig[0]) + RLOG(sig[1]) + RLOG(sig[2]);
    for (int i = 0; i < 3; i++) tau[i] = R(2.0) * mu * eps[i] + lam * sum * R(1.0);
    diag_sandwich(U, tau, V, t);
    mat_T(F, FT);
    mat_mul(t, FT, out);
}
/* mpm_utils.py:69-84 */
static void kirchoff_stress_drucker_prager(const real *F, const real *U, const real *V,
                                           const real *sig, real mu, real lam, real *out) {
    real sum = RLOG(sig[0]) + RLOG(sig[1]) + RLOG(sig[2]);
    real c[3], t[9], FT[9];
    for (int i = 0; i < 3; i++)
        c[i] = R(2.0) * mu * RLOG(sig[i]) * (R(1.0) / sig[i]) + lam * sum * (R(1.0) / sig[i]);
    diag_sandwich(U, c, V, t);
    mat_T(F, FT);
    mat_mul(t, FT, out);
}
/* mpm_utils.py:87-99 */
static void inverse_lower_triangle(const real *M, real *o) {
    real M11 = M[0], M21 = M[3], M22 = M[4], M31 = M[6], M32 = M[7], M33 = M[8];
    real invdet = R(1.0) / (M11 * M22 * M33);
    real t[9] = {M22 * M33, 0, 0, -M21 * M33, M11 * M33, 0, M21 * M32 - M31 * M22, -M11 * M32, M11 * M22};
    for (int i = 0; i < 9; i++) o[i] = invdet * t[i];
}

/* mpm_utils.py:101-177.  Returns stress (vol * P3 (x) d3) in out[9] and the three
 * vertex forces in f1,f2,f3 (the caller scatters them, :172-175). */
void orc_kirchoff_stress_anisotropy(const real *R_inv, const real *d, real vol, real mu, real lam,
                                    real gamma, real kappa, real *out, real *f1, real *f2, real *f3) {
    real iD11 = R_inv[0], iD12 = R_inv[1], iD22 = R_inv[2];
    real Q[9], Rm[9];
    orc_qr3_signed(d, Q, Rm);
    real F11 = Rm[0] * iD11;
    real F12 = Rm[0] * iD12 + Rm[1] * iD22;
    real F22 = Rm[4] * iD22;
    real RiDT[9] = {F11, 0, 0, F12, F22, 0, Rm[2], Rm[5], Rm[8]};
    real iFTJ[4] = {F22, 0, -F12, F11};
    /* svd3 of [[F11,F12,0],[0,F22,0],[0,0,0]] ; Rot = U2 V2^T (:133-141) */
    real F3[9] = {F11, F12, 0, 0, F22, 0, 0, 0, 0}, U3[9], V3[9], s3[3];
    orc_svd3(F3, U3, s3, V3);
    real Rot[4];
    Rot[0] = U3[0] * V3[0] + U3[1] * V3[1];
    Rot[1] = U3[0] * V3[3] + U3[1] * V3[4];
    Rot[2] = U3[3] * V3[0] + U3[4] * V3[1];
    Rot[3] = U3[3] * V3[3] + U3[4] * V3[4];
    real J = F11 * F22;
    real F2[4] = {F11, F12, 0, F22};
    real K2[4];
    for (int i = 0; i < 4; i++) K2[i] = R(2.0) * mu * (F2[i] - Rot[i]) + lam * (J - R(1.0)) * iFTJ[i];
    real dr11 = K2[0], dr12 = K2[1], dr22 = K2[3];
    real dr13 = gamma * Rm[2], dr23 = gamma * Rm[5], dr33;
    if (Rm[8] > R(1.0)) dr33 = R(0.0);
    else dr33 = -kappa * (R(1.0) - Rm[8]) * (R(1.0) - Rm[8]);
    real dr[9] = {dr11, dr12, dr13, 0, dr22, dr23, 0, 0, dr33};
    real K3[9];
    mat_mul(dr, RiDT, K3);
    real K3s[9] = {K3[0], K3[1], K3[2], K3[1], K3[4], K3[5], K3[2], K3[5], K3[8]};
    real RiDTinv[9], QK[9], P[9];
    inverse_lower_triangle(RiDT, RiDTinv);
    mat_mul(Q, K3s, QK);
    mat_mul(QK, RiDTinv, P);
    real P1[3] = {P[0], P[3], P[6]}, P2[3] = {P[1], P[4], P[7]}, P3[3] = {P[2], P[5], P[8]};
    real d3[3] = {d[2], d[5], d[8]};
    for (int i = 0; i < 3; i++) {
        f2[i] = -vol * (iD11 * P1[i] + iD12 * P2[i]);
        f3[i] = -vol * iD22 * P2[i];
        f1[i] = -(f2[i] + f3[i]);
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) out[3 * r + c] = vol * (P3[r] * d3[c]);
}

/* mpm_utils.py:179-209 */
void orc_anisotropy_return_mapping(const real *d, real kappa, real gamma, real friction_coeff,
                                   real *new_d) {
    real Q[9], R2[9], Rm[9];
    orc_qr3_signed(d, Q, R2);
    memcpy(Rm, R2, sizeof Rm);
    if (R2[8] > R(1.0)) {
        Rm[6] = 0; Rm[7] = 0; Rm[8] = R(1.0);
    } else {
        real fn = kappa * (R(1.0) - R2[8]) * (R(1.0) - R2[8]);
        real ff = gamma * RSQRT(R2[2] * R2[2] + R2[5] * R2[5]);
        if (ff > friction_coeff * fn) {
            Rm[2] = R2[2] * friction_coeff * fn / ff;
            Rm[5] = R2[5] * friction_coeff * fn / ff;
        }
    }
    real d3[3];
    for (int r = 0; r < 3; r++) d3[r] = Q[3 * r] * Rm[2] + Q[3 * r + 1] * Rm[5] + Q[3 * r + 2] * Rm[8];
    memcpy(new_d, d, 9 * sizeof(real));
    new_d[2] = d3[0]; new_d[5] = d3[1]; new_d[8] = d3[2];
}

/* mpm_utils.py:212-255 */
static void von_mises_return_mapping(const real *Ft, OrcSim *S, int p, real *Fo) {
    real U[9], V[9], so[3], sig[3], eps[3], tau[3];
    orc_svd3(Ft, U, so, V);
    for (int i = 0; i < 3; i++) sig[i] = so[i] > R(0.01) ? so[i] : R(0.01);
    for (int i = 0; i < 3; i++) eps[i] = RLOG(sig[i]);
    real temp = (eps[0] + eps[1] + eps[2]) / R(3.0);
    for (int i = 0; i < 3; i++) tau[i] = R(2.0) * S->mu[p] * eps[i] + S->lam[p] * (eps[0] + eps[1] + eps[2]) * R(1.0);
    real st = tau[0] + tau[1] + tau[2];
    real cond[3] = {tau[0] - st / R(3.0), tau[1] - st / R(3.0), tau[2] - st / R(3.0)};
    if (vlen(cond) > S->yield_stress[p]) {
        real eh[3] = {eps[0] - temp, eps[1] - temp, eps[2] - temp};
        real ehn = vlen(eh) + R(1e-6);
        real dg = ehn - S->yield_stress[p] / (R(2.0) * S->mu[p]);
        real se[3];
        for (int i = 0; i < 3; i++) { eps[i] = eps[i] - (dg / ehn) * eh[i]; se[i] = REXP(eps[i]); }
        diag_sandwich(U, se, V, Fo);
        if (S->hardening_i == 1)
            S->yield_stress[p] = S->yield_stress[p] + R(2.0) * S->mu[p] * S->xi * dg;
    } else memcpy(Fo, Ft, 9 * sizeof(real));
}
/* mpm_utils.py:258-311 */
static void von_mises_return_mapping_with_damage(const real *Ft, OrcSim *S, int p, real *Fo) {
    real U[9], V[9], so[3], sig[3], eps[3], tau[3];
    orc_svd3(Ft, U, so, V);
    for (int i = 0; i < 3; i++) sig[i] = so[i] > R(0.01) ? so[i] : R(0.01);
    for (int i = 0; i < 3; i++) eps[i] = RLOG(sig[i]);
    real temp = (eps[0] + eps[1] + eps[2]) / R(3.0);
    for (int i = 0; i < 3; i++) tau[i] = R(2.0) * S->mu[p] * eps[i] + S->lam[p] * (eps[0] + eps[1] + eps[2]) * R(1.0);
    real st = tau[0] + tau[1] + tau[2];
    real cond[3] = {tau[0] - st / R(3.0), tau[1] - st / R(3.0), tau[2] - st / R(3.0)};
    if (vlen(cond) > S->yield_stress[p]) {
        if (S->yield_stress[p] <= 0) { memcpy(Fo, Ft, 9 * sizeof(real)); return; }
        real eh[3] = {eps[0] - temp, eps[1] - temp, eps[2] - temp};
        real ehn = vlen(eh) + R(1e-6);
        real dg = ehn - S->yield_stress[p] / (R(2.0) * S->mu[p]);
        real se[3], corr[3];
        for (int i = 0; i < 3; i++) { corr[i] = (dg / ehn) * eh[i]; eps[i] = eps[i] - corr[i]; }
        S->yield_stress[p] = S->yield_stress[p] - S->softening * vlen(corr);
        if (S->yield_stress[p] <= 0) { S->mu[p] = R(0.0); S->lam[p] = R(0.0); }
        for (int i = 0; i < 3; i++) se[i] = REXP(eps[i]);
        diag_sandwich(U, se, V, Fo);
        if (S->hardening_i == 1)
            S->yield_stress[p] = S->yield_stress[p] + R(2.0) * S->mu[p] * S->xi * dg;
    } else memcpy(Fo, Ft, 9 * sizeof(real));
}
/* mpm_utils.py:315-359 */
static void viscoplasticity_return_mapping_with_StVK(const real *Ft, OrcSim *S, int p, real dt, real *Fo) {
    real U[9], V[9], so[3], sig[3], eps[3], b[3];
    orc_svd3(Ft, U, so, V);
    for (int i = 0; i < 3; i++) sig[i] = so[i] > R(0.01) ? so[i] : R(0.01);
    for (int i = 0; i < 3; i++) { b[i] = sig[i] * sig[i]; eps[i] = RLOG(sig[i]); }
    real tr = eps[0] + eps[1] + eps[2];
    real eh[3] = {eps[0] - tr / R(3.0), eps[1] - tr / R(3.0), eps[2] - tr / R(3.0)};
    real st[3] = {R(2.0) * S->mu[p] * eh[0], R(2.0) * S->mu[p] * eh[1], R(2.0) * S->mu[p] * eh[2]};
    real stn = vlen(st);
    real y = stn - RSQRT(R(2.0) / R(3.0)) * S->yield_stress[p];
    if (y > 0) {
        real mu_hat = S->mu[p] * (b[0] + b[1] + b[2]) / R(3.0);
        real snn = stn - y / (R(1.0) + S->plastic_viscosity / (R(2.0) * mu_hat * dt));
        real se[3];
        for (int i = 0; i < 3; i++) {
            real sn = (snn / stn) * st[i];
            real en = R(1.0) / (R(2.0) * S->mu[p]) * sn + tr / R(3.0);
            se[i] = REXP(en);
        }
        diag_sandwich(U, se, V, Fo);
    } else memcpy(Fo, Ft, 9 * sizeof(real));
}
/* mpm_utils.py:362-399 */
static void sand_return_mapping(const real *Ft, OrcSim *S, int p, real *Fo) {
    real U[9], V[9], sig[3], eps[3];
    orc_svd3(Ft, U, sig, V);
    for (int i = 0; i < 3; i++) {
        real a = RABS(sig[i]);
        eps[i] = RLOG(a > R(1e-14) ? a : R(1e-14));
    }
    real tr = eps[0] + eps[1] + eps[2];
    real eh[3] = {eps[0] - tr / R(3.0), eps[1] - tr / R(3.0), eps[2] - tr / R(3.0)};
    real ehn = vlen(eh);
    real dg = ehn + (R(3.0) * S->lam[p] + R(2.0) * S->mu[p]) / (R(2.0) * S->mu[p]) * tr * S->alpha;
    memcpy(Fo, Ft, 9 * sizeof(real)); /* delta_gamma <= 0 */
    if (dg > 0 && tr > 0) {
        real VT[9];
        mat_T(V, VT);
        mat_mul(U, VT, Fo);
    }
    if (dg > 0 && tr <= 0) {
        real s[3];
        for (int i = 0; i < 3; i++) s[i] = REXP(eps[i] - eh[i] * (dg / ehn));
        diag_sandwich(U, s, V, Fo);
    }
}

/* ------------------------------------------------------------- grid kernels */
/* mpm_utils.py:411-417 */
void orc_zero_grid(OrcSim *S) {
    size_t n3 = (size_t)S->n_grid * S->n_grid * S->n_grid;
    memset(S->grid_m, 0, n3 * sizeof(real));
    memset(S->grid_v_in, 0, 3 * n3 * sizeof(real));
    memset(S->grid_v_out, 0, 3 * n3 * sizeof(real));
}

static inline void atomic_add(real *p, real v) {
#ifdef _OPENMP
#pragma omp atomic
#endif
    *p += v;
}

/* quadratic B-spline stencil shared by P2G / G2P / collider / mover
 * (mpm_utils.py:499-514).  w[a][i]: axis a, stencil offset i (wp.mat33 columns). */
static inline void stencil(const real *xp, real inv_dx, int *base, real *fx, real w[3][3], real dw[3][3]) {
    for (int a = 0; a < 3; a++) {
        real gp = xp[a] * inv_dx;
        base[a] = (int)(gp - R(0.5)); /* wp.int truncates toward zero */
        fx[a] = gp - (real)base[a];
        real wa = R(1.5) - fx[a], wb = fx[a] - R(1.0), wc = fx[a] - R(0.5);
        w[a][0] = wa * wa * R(0.5);
        w[a][1] = R(0.0) - wb * wb + R(0.75);
        w[a][2] = wc * wc * R(0.5);
        dw[a][0] = fx[a] - R(1.5);
        dw[a][1] = R(-2.0) * (fx[a] - R(1.0));
        dw[a][2] = fx[a] - R(0.5);
    }
}
#define GIDX(S, ix, iy, iz) ((((size_t)(ix)) * (S)->n_grid + (iy)) * (S)->n_grid + (iz))
static inline int in_grid(const OrcSim *S, int ix, int iy, int iz) {
    return ix >= 0 && iy >= 0 && iz >= 0 && ix < S->n_grid && iy < S->n_grid && iz < S->n_grid;
}

/* mpm_utils.py:1017-1105, launched with dim n_no_vertices (mpm_solver.py:327-332) */
void orc_compute_stress_from_F_trial(OrcSim *S, real dt) {
    int nnv = S->n_particles - S->n_vertices;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int p = 0; p < nnv; p++) {
        if (S->selection[p] != 0) continue;
        real stress[9] = {0};
        if (S->elements[p] == 1) {
            real nd[9], f1[3], f2[3], f3[3];
            orc_anisotropy_return_mapping(&S->d[9 * p], S->kappa[p], S->gamma[p], S->friction_coeff, nd);
            memcpy(&S->d[9 * p], nd, sizeof nd);
            orc_kirchoff_stress_anisotropy(&S->R_inv[3 * p], &S->d[9 * p], S->vol[p], S->mu[p], S->lam[p],
                                           S->gamma[p], S->kappa[p], stress, f1, f2, f3);
            int v1 = (int)S->faces[3 * p], v2 = (int)S->faces[3 * p + 1], v3 = (int)S->faces[3 * p + 2];
            for (int i = 0; i < 3; i++) {
                atomic_add(&S->vertex_force[3 * v1 + i], f1[i]);
                atomic_add(&S->vertex_force[3 * v2 + i], f2[i]);
                atomic_add(&S->vertex_force[3 * v3 + i], f3[i]);
            }
        } else if (S->traditional[p] == 1) {
            real *F = &S->F[9 * p];
            const real *Ft = &S->F_trial[9 * p];
            if (S->material == 1) von_mises_return_mapping(Ft, S, p, F);
            else if (S->material == 2) sand_return_mapping(Ft, S, p, F);
            else if (S->material == 3) viscoplasticity_return_mapping_with_StVK(Ft, S, p, dt, F);
            else if (S->material == 5) von_mises_return_mapping_with_damage(Ft, S, p, F);
            else memcpy(F, Ft, 9 * sizeof(real));
            real J = mat_det(F), U[9], V[9], sig[3];
            orc_svd3(F, U, sig, V);
            if (S->material == 0 || S->material == 5) kirchoff_stress_FCR(F, U, V, J, S->mu[p], S->lam[p], stress);
            if (S->material == 1) kirchoff_stress_StVK(F, U, V, sig, S->mu[p], S->lam[p], stress);
            if (S->material == 2) kirchoff_stress_drucker_prager(F, U, V, sig, S->mu[p], S->lam[p], stress);
            if (S->material == 3) kirchoff_stress_StVK(F, U, V, sig, S->mu[p], S->lam[p], stress);
            real sT[9];
            mat_T(stress, sT);
            for (int i = 0; i < 9; i++) stress[i] = (stress[i] + sT[i]) / R(2.0);
        }
        memcpy(&S->stress[9 * p], stress, sizeof stress);
    }
}

/* mpm_utils.py:484-557; offset = n_no_vertices (mpm_solver.py:358) */
void orc_p2g_apic_with_stress(OrcSim *S, real dt) {
    int offset = S->n_particles - S->n_vertices;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int p = 0; p < S->n_particles; p++) {
        if (S->selection[p] != 0) continue;
        real vforce[3] = {0, 0, 0}, stress[9] = {0};
        if (S->vertices[p] == 1) {
            for (int i = 0; i < 3; i++) vforce[i] = S->vertex_force[3 * (p - offset) + i];
        } else if (S->traditional[p] == 1) {
            for (int i = 0; i < 9; i++) stress[i] = S->vol[p] * S->stress[9 * p + i];
        } else {
            for (int i = 0; i < 9; i++) stress[i] = S->stress[9 * p + i];
        }
        int base[3];
        real fx[3], w[3][3], dw[3][3];
        stencil(&S->x[3 * p], S->inv_dx, base, fx, w, dw);
        real C[9];
        const real *Cp = &S->C[9 * p];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                C[3 * r + c] = (R(1.0) - S->rpic_damping) * Cp[3 * r + c] +
                               S->rpic_damping / R(2.0) * (Cp[3 * r + c] - Cp[3 * c + r]);
        if (S->rpic_damping < R(-0.001)) memset(C, 0, sizeof C);
        const real *vp = &S->v[3 * p];
        real m = S->mass[p];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) {
                    real dpos[3] = {((real)i - fx[0]) * S->dx, ((real)j - fx[1]) * S->dx, ((real)k - fx[2]) * S->dx};
                    int ix = base[0] + i, iy = base[1] + j, iz = base[2] + k;
                    real weight = w[0][i] * w[1][j] * w[2][k];
                    real dwt[3] = {dw[0][i] * w[1][j] * w[2][k] * S->inv_dx, w[0][i] * dw[1][j] * w[2][k] * S->inv_dx,
                                   w[0][i] * w[1][j] * dw[2][k] * S->inv_dx};
                    real force[3];
                    if (S->vertices[p] == 1) {
                        for (int a = 0; a < 3; a++) force[a] = weight * vforce[a];
                    } else {
                        for (int a = 0; a < 3; a++)
                            force[a] = -(stress[3 * a] * dwt[0] + stress[3 * a + 1] * dwt[1] + stress[3 * a + 2] * dwt[2]);
                    }
                    if (!in_grid(S, ix, iy, iz)) continue; /* reference has no check; guard only */
                    size_t g = GIDX(S, ix, iy, iz);
                    for (int a = 0; a < 3; a++) {
                        real Cd = C[3 * a] * dpos[0] + C[3 * a + 1] * dpos[1] + C[3 * a + 2] * dpos[2];
                        real add = weight * m * (vp[a] + Cd) + dt * force[a];
                        atomic_add(&S->grid_v_in[3 * g + a], add);
                    }
                    atomic_add(&S->grid_m[g], weight * m);
                }
    }
}

/* mpm_utils.py:561-572 */
void orc_grid_normalization_and_gravity(OrcSim *S, real dt) {
    size_t n3 = (size_t)S->n_grid * S->n_grid * S->n_grid;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t g = 0; g < n3; g++) {
        if (S->grid_m[g] > R(1e-15)) {
            real inv = R(1.0) / S->grid_m[g];
            for (int a = 0; a < 3; a++) S->grid_v_out[3 * g + a] = S->grid_v_in[3 * g + a] * inv + dt * S->g[a];
        }
    }
}
/* mpm_utils.py:1162-1174 */
void orc_add_damping_via_grid(OrcSim *S, real scale) {
    size_t n3 = (size_t)S->n_grid * S->n_grid * S->n_grid;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t g = 0; g < 3 * n3; g++) S->grid_v_out[g] -= (R(1.0) - scale) * S->grid_v_out[g];
}

/* scatter helper for collider / mover: bounds check 0<=base<dim-3 (mpm_solver.py:692,858) */
static inline int scatter_ok(const OrcSim *S, const int *b) {
    int n = S->n_grid;
    return b[0] >= 0 && b[0] < n - 3 && b[1] >= 0 && b[1] < n - 3 && b[2] >= 0 && b[2] < n - 3;
}

/* mpm_solver.py:819-917: zero_grid, compute_mesh, normalize_grid, collide */
void orc_mesh_collider(OrcSim *S) {
    size_t n3 = (size_t)S->n_grid * S->n_grid * S->n_grid;
    memset(S->col_weight, 0, n3 * sizeof(real));
    memset(S->col_v_in, 0, 3 * n3 * sizeof(real));
    memset(S->col_v_out, 0, 3 * n3 * sizeof(real));
    memset(S->col_normal, 0, 3 * n3 * sizeof(real));
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int f = 0; f < S->n_mesh_f; f++) {
        int i0 = S->mesh_faces[3 * f], i1 = S->mesh_faces[3 * f + 1], i2 = S->mesh_faces[3 * f + 2];
        const real *p0 = &S->mesh_points[3 * i0], *p1 = &S->mesh_points[3 * i1], *p2 = &S->mesh_points[3 * i2];
        const real *v0 = &S->mesh_velocities[3 * i0], *v1 = &S->mesh_velocities[3 * i1], *v2 = &S->mesh_velocities[3 * i2];
        real fp[3], fv[3], e1[3], e2[3], cr[3], fn[3];
        for (int a = 0; a < 3; a++) {
            fp[a] = (p0[a] + p1[a] + p2[a]) / R(3.0);
            fv[a] = (v0[a] + v1[a] + v2[a]) / R(3.0);
            e1[a] = p1[a] - p0[a];
            e2[a] = p2[a] - p0[a];
        }
        cr[0] = e1[1] * e2[2] - e1[2] * e2[1];
        cr[1] = e1[2] * e2[0] - e1[0] * e2[2];
        cr[2] = e1[0] * e2[1] - e1[1] * e2[0];
        vnormalize(cr, fn);
        int base[3];
        real fx[3], w[3][3], dw[3][3];
        stencil(fp, S->inv_dx, base, fx, w, dw);
        if (!scatter_ok(S, base)) continue;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) {
                    size_t g = GIDX(S, base[0] + i, base[1] + j, base[2] + k);
                    real weight = w[0][i] * w[1][j] * w[2][k];
                    for (int a = 0; a < 3; a++) {
                        atomic_add(&S->col_v_in[3 * g + a], weight * fv[a]);
                        atomic_add(&S->col_normal[3 * g + a], weight * fn[a]);
                    }
                    atomic_add(&S->col_weight[g], weight);
                }
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t g = 0; g < n3; g++) {
        if (S->col_weight[g] > R(1e-15)) {
            real inv = R(1.0) / S->col_weight[g];
            for (int a = 0; a < 3; a++) S->col_v_out[3 * g + a] = S->col_v_in[3 * g + a] * inv;
        }
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t g = 0; g < n3; g++) {
        real *v = &S->grid_v_out[3 * g];
        if (S->col_weight[g] > R(1e-15)) {
            const real *mv = &S->col_v_out[3 * g];
            real vrel[3] = {v[0] - mv[0], v[1] - mv[1], v[2] - mv[2]}, n[3], vproj[3], vfric[3];
            vnormalize(&S->col_normal[3 * g], n);
            real nc = vrel[0] * n[0] + vrel[1] * n[1] + vrel[2] * n[2];
            real mn = nc < R(0.0) ? nc : R(0.0);
            for (int a = 0; a < 3; a++) vproj[a] = vrel[a] - mn * n[a];
            real lp = vlen(vproj);
            if (nc < R(0.0) && lp > R(1e-20)) {
                real s = lp + nc * S->collider_friction;
                if (s < R(0.0)) s = R(0.0);
                real nv[3];
                vnormalize(vproj, nv);
                for (int a = 0; a < 3; a++) vfric[a] = s * nv[a];
            } else {
                for (int a = 0; a < 3; a++) vfric[a] = vproj[a];
            }
            for (int a = 0; a < 3; a++) v[a] = vfric[a] + mv[a];
        }
    }
}

/* mpm_solver.py:669-799.  joint_t_v: prescribed velocities for the LAST n_joint_t
 * traditional particles (offset n_no_vertices - n_joint_t, :446); joint_v_v for the
 * first num_joint_v vertex particles (offset n_no_vertices, :458); joint_f_v for the
 * first num_joint_f element particles (:462-472). */
static void mover_scatter(OrcSim *S, const real *xp, const real *vel) {
    int base[3];
    real fx[3], w[3][3], dw[3][3];
    stencil(xp, S->inv_dx, base, fx, w, dw);
    if (!scatter_ok(S, base)) return;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                size_t g = GIDX(S, base[0] + i, base[1] + j, base[2] + k);
                real weight = w[0][i] * w[1][j] * w[2][k];
                for (int a = 0; a < 3; a++) atomic_add(&S->mov_velocity[3 * g + a], weight * vel[a]);
                atomic_add(&S->mov_weight[g], weight);
            }
}
void orc_particle_mover(OrcSim *S, const real *joint_t_v, int n_joint_t, const real *joint_v_v,
                        const real *joint_f_v) {
    size_t n3 = (size_t)S->n_grid * S->n_grid * S->n_grid;
    int nnv = S->n_particles - S->n_vertices;
    memset(S->mov_weight, 0, n3 * sizeof(real));
    memset(S->mov_velocity, 0, 3 * n3 * sizeof(real));
    if (joint_t_v)
        for (int p = 0; p < n_joint_t; p++) mover_scatter(S, &S->x[3 * (p + nnv - n_joint_t)], &joint_t_v[3 * p]);
    for (int p = 0; p < S->num_joint_v; p++) mover_scatter(S, &S->x[3 * (p + nnv)], &joint_v_v[3 * p]);
    for (int p = 0; p < S->num_joint_f; p++) mover_scatter(S, &S->x[3 * p], &joint_f_v[3 * p]);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t g = 0; g < n3; g++) {
        if (S->mov_weight[g] > R(1e-15)) {
            real inv = R(1.0) / S->mov_weight[g];
            for (int a = 0; a < 3; a++) S->grid_v_out[3 * g + a] = S->mov_velocity[3 * g + a] * inv;
        }
    }
}

/* grid_postprocess kernels: surface (mpm_solver.py:600-655), cuboid (:950-981),
 * bounding box (:993-1050), mask (:1341-1352) */
void orc_apply_bc(OrcSim *S, OrcBC *bc, real time, real dt) {
    int n = S->n_grid;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int gx = 0; gx < n; gx++)
        for (int gy = 0; gy < n; gy++)
            for (int gz = 0; gz < n; gz++) {
                real *v = &S->grid_v_out[3 * GIDX(S, gx, gy, gz)];
                if (bc->kind == BC_SURFACE) {
                    if (time >= bc->start_time && time < bc->end_time) {
                        real off[3] = {(real)gx * S->dx - bc->point[0], (real)gy * S->dx - bc->point[1],
                                       (real)gz * S->dx - bc->point[2]};
                        real dp = off[0] * bc->normal[0] + off[1] * bc->normal[1] + off[2] * bc->normal[2];
                        if (dp < R(0.0)) {
                            if (bc->surface_type == 0) {
                                v[0] = v[1] = v[2] = R(0.0);
                            } else if (bc->surface_type == 11) {
                                if ((real)gz * S->dx < R(0.4) || (real)gz * S->dx > R(0.53)) {
                                    v[0] = v[1] = v[2] = R(0.0);
                                } else {
                                    v[0] = v[0] * R(0.3); v[1] = R(0.0) * R(0.3); v[2] = v[2] * R(0.3);
                                }
                            } else {
                                /* slip / frictional branches compute a velocity and then
                                 * store zero anyway (mpm_solver.py:636-655) */
                                v[0] = v[1] = v[2] = R(0.0);
                            }
                        }
                    }
                } else if (bc->kind == BC_CUBOID) {
                    if (time >= bc->start_time && time < bc->end_time) {
                        real off[3] = {(real)gx * S->dx - bc->point[0], (real)gy * S->dx - bc->point[1],
                                       (real)gz * S->dx - bc->point[2]};
                        if (RABS(off[0]) < bc->size[0] && RABS(off[1]) < bc->size[1] && RABS(off[2]) < bc->size[2]) {
                            v[0] = bc->velocity[0]; v[1] = bc->velocity[1]; v[2] = bc->velocity[2];
                        }
                    } else if (bc->reset == 1) {
                        if (time < bc->end_time + R(15.0) * dt) v[0] = v[1] = v[2] = R(0.0);
                    }
                } else if (bc->kind == BC_BBOX) {
                    int pad = 3;
                    if (time >= bc->start_time && time < bc->end_time) {
                        if (gx < pad && v[0] < 0) v[0] = R(0.0);
                        if (gx >= n - pad && v[0] > 0) v[0] = R(0.0);
                        if (gy < pad && v[1] < 0) v[1] = R(0.0);
                        if (gy >= n - pad && v[1] > 0) v[1] = R(0.0);
                        if (gz < pad && v[2] < 0) v[2] = R(0.0);
                        if (gz >= n - pad && v[2] > 0) v[2] = R(0.0);
                    }
                } else if (bc->kind == BC_MASK) {
                    if (bc->mask[GIDX(S, gx, gy, gz)] >= 1) v[0] = v[1] = v[2] = R(0.0);
                }
            }
    /* host-side modify_bc of the moving cuboid (mpm_solver.py:975-981) */
    if (bc->kind == BC_CUBOID && time >= bc->start_time && time < bc->end_time)
        for (int a = 0; a < 3; a++) bc->point[a] = bc->point[a] + dt * bc->velocity[a];
}

/* shared gather of g2p_v / g2p_e (mpm_utils.py:726-763, 798-836) */
static void g2p_gather(const OrcSim *S, const real *xp, real *new_v, real *new_C, real *new_F) {
    int base[3];
    real fx[3], w[3][3], dw[3][3];
    stencil(xp, S->inv_dx, base, fx, w, dw);
    for (int a = 0; a < 3; a++) new_v[a] = 0;
    for (int a = 0; a < 9; a++) { new_C[a] = 0; new_F[a] = 0; }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                int ix = base[0] + i, iy = base[1] + j, iz = base[2] + k;
                if (!in_grid(S, ix, iy, iz)) continue; /* guard only */
                real dpos[3] = {(real)i - fx[0], (real)j - fx[1], (real)k - fx[2]};
                real weight = w[0][i] * w[1][j] * w[2][k];
                const real *gv = &S->grid_v_out[3 * GIDX(S, ix, iy, iz)];
                real dwt[3] = {dw[0][i] * w[1][j] * w[2][k] * S->inv_dx, w[0][i] * dw[1][j] * w[2][k] * S->inv_dx,
                               w[0][i] * w[1][j] * dw[2][k] * S->inv_dx};
                real s = weight * S->inv_dx * R(4.0);
                for (int r = 0; r < 3; r++) {
                    new_v[r] = new_v[r] + gv[r] * weight;
                    for (int c = 0; c < 3; c++) {
                        new_C[3 * r + c] = new_C[3 * r + c] + (gv[r] * dpos[c]) * s;
                        new_F[3 * r + c] = new_F[3 * r + c] + gv[r] * dwt[c];
                    }
                }
            }
}
static inline real clampr(real x, real a, real b) { return x < a ? a : (x > b ? b : x); }

/* mpm_utils.py:716-786; dim N-Ne, offset Ne (mpm_solver.py:518-523) */
void orc_g2p_v(OrcSim *S, real dt) {
    int offset = S->n_elements, cnt = S->n_particles - S->n_elements;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int q = 0; q < cnt; q++) {
        int p = q + offset;
        if (S->selection[p] != 0) continue;
        real nv[3], nC[9], nF[9];
        g2p_gather(S, &S->x[3 * p], nv, nC, nF);
        real dx = R(1.0) / S->inv_dx, a_min = dx * R(2.0), a_max = S->grid_lim - dx * R(2.0);
        for (int a = 0; a < 3; a++) {
            S->v[3 * p + a] = nv[a];
            S->x[3 * p + a] = clampr(S->x[3 * p + a] + dt * nv[a], a_min, a_max);
        }
        memcpy(&S->C[9 * p], nC, sizeof nC);
        if (S->traditional[p] == 1) {
            real M[9], o[9];
            for (int i = 0; i < 9; i++) M[i] = nF[i] * dt;
            M[0] += R(1.0); M[4] += R(1.0); M[8] += R(1.0);
            mat_mul(M, &S->F[9 * p], o);
            memcpy(&S->F_trial[9 * p], o, sizeof o);
        }
    }
}
/* mpm_utils.py:788-857; dim Ne, offset n_no_vertices (mpm_solver.py:529-534) */
void orc_g2p_e(OrcSim *S, real dt) {
    int offset = S->n_particles - S->n_vertices;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int p = 0; p < S->n_elements; p++) {
        if (S->selection[p] != 0) continue;
        real nv[3], nC[9], nF[9];
        g2p_gather(S, &S->x[3 * p], nv, nC, nF);
        int v1 = (int)S->faces[3 * p] + offset, v2 = (int)S->faces[3 * p + 1] + offset,
            v3 = (int)S->faces[3 * p + 2] + offset;
        for (int a = 0; a < 3; a++) {
            S->v[3 * p + a] = (S->v[3 * v1 + a] + S->v[3 * v2 + a] + S->v[3 * v3 + a]) / R(3.0);
            S->x[3 * p + a] = (S->x[3 * v1 + a] + S->x[3 * v2 + a] + S->x[3 * v3 + a]) / R(3.0);
        }
        memcpy(&S->C[9 * p], nC, sizeof nC);
        real *d = &S->d[9 * p];
        real d3[3] = {d[2], d[5], d[8]}, d3t[3];
        for (int r = 0; r < 3; r++) {
            real m0 = nF[3 * r] * dt + (r == 0 ? R(1.0) : R(0.0));
            real m1 = nF[3 * r + 1] * dt + (r == 1 ? R(1.0) : R(0.0));
            real m2 = nF[3 * r + 2] * dt + (r == 2 ? R(1.0) : R(0.0));
            d3t[r] = m0 * d3[0] + m1 * d3[1] + m2 * d3[2];
        }
        for (int r = 0; r < 3; r++) {
            d[3 * r] = S->x[3 * v2 + r] - S->x[3 * v1 + r];
            d[3 * r + 1] = S->x[3 * v3 + r] - S->x[3 * v1 + r];
            d[3 * r + 2] = d3t[r];
        }
    }
}

/* One full substep in the reference's launch order (mpm_solver.py:229-536).
 * mesh_x / mesh_v: new body-mesh points / velocities or NULL (:285-315).
 * The mover runs only when both joint_v_v and joint_f_v are given (:421). */
void orc_p2g2p(OrcSim *S, real dt, real time, const real *mesh_x, const real *mesh_v, const real *joint_t_v,
               int n_joint_t, const real *joint_v_v, const real *joint_f_v) {
    orc_zero_grid(S);
    memset(S->vertex_force, 0, (size_t)3 * S->n_vertices * sizeof(real));
    if (mesh_x) memcpy(S->mesh_points, mesh_x, (size_t)3 * S->n_mesh_v * sizeof(real));
    if (mesh_v) memcpy(S->mesh_velocities, mesh_v, (size_t)3 * S->n_mesh_v * sizeof(real));
    orc_compute_stress_from_F_trial(S, dt);
    orc_p2g_apic_with_stress(S, dt);
    orc_grid_normalization_and_gravity(S, dt);
    if (S->grid_v_damping_scale < R(1.0)) orc_add_damping_via_grid(S, S->grid_v_damping_scale);
    if (S->has_collider) orc_mesh_collider(S);
    if (S->has_mover && joint_v_v && joint_f_v) orc_particle_mover(S, joint_t_v, n_joint_t, joint_v_v, joint_f_v);
    for (int k = 0; k < S->n_bc; k++) orc_apply_bc(S, &S->bc[k], time, dt);
    orc_g2p_v(S, dt);
    orc_g2p_e(S, dt);
}

int orc_sizeof_real(void) { return (int)sizeof(real); }
int orc_sizeof_sim(void) { return (int)sizeof(OrcSim); }
int orc_sizeof_bc(void) { return (int)sizeof(OrcBC); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
