// Kernels of one MPM substep (launch order = MPMWARP.p2g2p, warp_mpm/mpm_solver.py:229-536).
//
// Every particle kernel is "slab staged": a warp owns 32 consecutive (cell-sorted) particles, pulls
// the sub-records it needs into shared memory with one cp.async.bulk each (TMA), works lane = particle
// (or lane = stencil node in P2G) out of shared memory, and pushes whole updated sub-records back with
// one bulk store each.  Global traffic is therefore exactly the records, fully coalesced, and the LSU
// only sees the irregular accesses (grid nodes, corner vertices).  Warps never synchronise with
// each other: one mbarrier per warp, no __syncthreads.
#pragma once
#include "mpm_device.cuh"

namespace mpm {

// ---- per-warp slab context -------------------------------------------------------------------
struct Warp {
    int lane, p0, cnt;
    uint64_t* bar;
    unsigned char* buf;  // this warp's shared-memory region
};
// NW warps per CTA, WB bytes of shared memory per warp (after a 128-byte barrier header)
template <int NW, int WB>
__device__ __forceinline__ bool warp_begin(Warp& w, int n, unsigned char* smem, int blk = -1) {
    const int warp = threadIdx.x >> 5;
    w.lane = threadIdx.x & 31;
    w.p0 = ((blk < 0 ? (int)blockIdx.x : blk) * NW + warp) * 32;
    if (w.p0 >= n) return false;
    w.cnt = min(32, n - w.p0);
    w.bar = reinterpret_cast<uint64_t*>(smem) + warp;
    w.buf = smem + 128 + warp * WB;
    if (w.lane == 0) mbar_init(w.bar, 1);
    __syncwarp();
    return true;
}
struct Slab {
    void* s;        // shared-memory destination / source
    const void* g;  // base of the global sub-record array
    int F;          // 4-byte words per record
};
// one cp.async.bulk per sub-record slab, all completing on the warp's mbarrier
template <int N>
__device__ __forceinline__ void slab_issue(const Warp& w, const Slab (&sl)[N]) {
    if (w.lane == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < N; i++) tot += slab_bytes(w.cnt, sl[i].F);
        mbar_expect_tx(w.bar, tot);
#pragma unroll
        for (int i = 0; i < N; i++)
            bulk_g2s(sl[i].s, static_cast<const float*>(sl[i].g) + (size_t)w.p0 * sl[i].F, slab_bytes(w.cnt, sl[i].F), w.bar);
    }
}
__device__ __forceinline__ void slab_wait(const Warp& w) {
    while (!mbar_try_wait(w.bar, 0)) {}
}
template <int N>
__device__ __forceinline__ void slab_load(const Warp& w, const Slab (&sl)[N]) {
    slab_issue(w, sl);
    slab_wait(w);
}
template <int N>
__device__ __forceinline__ void slab_store(const Warp& w, const Slab (&sl)[N]) {
    fence_async_smem();
    __syncwarp();
    if (w.lane == 0) {
#pragma unroll
        for (int i = 0; i < N; i++)
            bulk_s2g(const_cast<float*>(static_cast<const float*>(sl[i].g)) + (size_t)w.p0 * sl[i].F, sl[i].s, slab_bytes(w.cnt, sl[i].F));
        bulk_commit_wait();
    }
}

// ============================================================ constitutive update (cloth elements)
// Fused anisotropy_return_mapping + kirchoff_stress_Anisotropy (mpm_utils.py:179-209, 101-177).
// The reference runs wp.qr3 twice on the same d1,d2; the second QR only differs in the third
// column of R, which is exactly the return-mapped (R02,R12,R22), so one QR serves both.
// wp.svd3 of [[F11,F12,0],[0,F22,0],[0,0,0]] is only used for U2 V2^T = polar rotation of the
// upper-triangular 2x2, which has the closed form [[a, b],[-b, a]]/|.|, a=F11+F22, b=F12.
struct ElemConst {
    float iD11, iD12, iD22, mu, lam, gamma, kappa, vol;
};
struct ElemStress {
    float nd3[3];                // return-mapped d3 (mpm_utils.py:205-208)
    float f1[3], f2[3], f3[3];   // corner forces (mpm_utils.py:163-175)
    float P3[3];                 // stress = vol * P3 (x) nd3 (mpm_utils.py:177)
};
// RETURN_MAP = false evaluates the stress of a d whose third column is ALREADY return mapped (what
// kirchoff_stress_Anisotropy sees in the reference); the mapping must not be applied twice: its R22 > 1
// branch keeps the shear while a second pass through the cone branch (fn = 0 at R22 = 1) would zero it.
template <bool RETURN_MAP = true>
__device__ __forceinline__ void element_stress(const float* d1, const float* d2, const float* d3, const ElemConst& k,
                                               float friction_coeff, ElemStress& o) {
    // rotation QR, sign-normalised (R00>0, R11>0, det Q=+1): Gram-Schmidt with q3 = q1 x q2, in
    // un-contracted IEEE arithmetic (see dot3_rn)
    const float r00 = len3_rn(d1);
    const float q1[3] = {__fdiv_rn(d1[0], r00), __fdiv_rn(d1[1], r00), __fdiv_rn(d1[2], r00)};
    const float r01 = dot3_rn(q1, d2);
    const float u2[3] = {__fsub_rn(d2[0], __fmul_rn(r01, q1[0])), __fsub_rn(d2[1], __fmul_rn(r01, q1[1])),
                         __fsub_rn(d2[2], __fmul_rn(r01, q1[2]))};
    const float r11 = len3_rn(u2);
    const float q2[3] = {__fdiv_rn(u2[0], r11), __fdiv_rn(u2[1], r11), __fdiv_rn(u2[2], r11)};
    const float q3[3] = {__fsub_rn(__fmul_rn(q1[1], q2[2]), __fmul_rn(q1[2], q2[1])),
                         __fsub_rn(__fmul_rn(q1[2], q2[0]), __fmul_rn(q1[0], q2[2])),
                         __fsub_rn(__fmul_rn(q1[0], q2[1]), __fmul_rn(q1[1], q2[0]))};
    float r02 = dot3_rn(q1, d3), r12 = dot3_rn(q2, d3), r22 = dot3_rn(q3, d3);
    // return mapping (mpm_utils.py:196-204)
    if (!RETURN_MAP) {
    } else if (r22 > 1.0f) {
        r22 = 1.0f;
    } else {
        float fn = k.kappa * (1.0f - r22) * (1.0f - r22);
        float ff = k.gamma * sqrtf(r02 * r02 + r12 * r12);
        if (ff > friction_coeff * fn) {
            float sc = friction_coeff * fn / ff;
            r02 *= sc;
            r12 *= sc;
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) o.nd3[r] = q1[r] * r02 + q2[r] * r12 + q3[r] * r22;
    // stress (mpm_utils.py:125-177) with R = [r00 r01 r02; 0 r11 r12; 0 0 r22]
    const float F11 = r00 * k.iD11, F12 = r00 * k.iD12 + r01 * k.iD22, F22 = r11 * k.iD22;
    const float pa = F11 + F22, pb = F12;
    const float pin = rsqrtf(pa * pa + pb * pb);
    const float c = pa * pin, s = pb * pin;  // Rot = [[c, s], [-s, c]]
    const float J = F11 * F22;
    const float lj = k.lam * (J - 1.0f);
    const float k00 = 2.0f * k.mu * (F11 - c) + lj * F22;
    const float k01 = 2.0f * k.mu * (F12 - s);
    const float k11 = 2.0f * k.mu * (F22 - c) + lj * F11;  // K2[1,0] is never used (mpm_utils.py:146-148)
    const float dr13 = k.gamma * r02, dr23 = k.gamma * r12;
    const float dr33 = (r22 > 1.0f) ? 0.0f : -k.kappa * (1.0f - r22) * (1.0f - r22);
    // K3 = dr * RiDT, RiDT = [F11 0 0; F12 F22 0; r02 r12 r22]
    const float K00 = k00 * F11 + k01 * F12 + dr13 * r02;
    const float K01 = k01 * F22 + dr13 * r12;
    const float K02 = dr13 * r22;
    const float K11 = k11 * F22 + dr23 * r12;
    const float K12 = dr23 * r22;
    const float K22 = dr33 * r22;
    // inverse of lower-triangular RiDT (mpm_utils.py:87-99)
    const float invdet = 1.0f / (F11 * F22 * r22);
    const float I00 = F22 * r22 * invdet, I10 = -F12 * r22 * invdet, I11 = F11 * r22 * invdet;
    const float I20 = (F12 * r12 - r02 * F22) * invdet, I21 = -F11 * r12 * invdet, I22 = F11 * F22 * invdet;
    // M = K3sym * RiDT^-1
    const float M00 = K00 * I00 + K01 * I10 + K02 * I20, M01 = K01 * I11 + K02 * I21, M02 = K02 * I22;
    const float M10 = K01 * I00 + K11 * I10 + K12 * I20, M11 = K11 * I11 + K12 * I21, M12 = K12 * I22;
    const float M20 = K02 * I00 + K12 * I10 + K22 * I20, M21 = K12 * I11 + K22 * I21, M22 = K22 * I22;
#pragma unroll
    for (int r = 0; r < 3; r++) {  // columns of P = Q M
        const float P1 = q1[r] * M00 + q2[r] * M10 + q3[r] * M20;
        const float P2 = q1[r] * M01 + q2[r] * M11 + q3[r] * M21;
        o.P3[r] = q1[r] * M02 + q2[r] * M12 + q3[r] * M22;
        o.f2[r] = -k.vol * (k.iD11 * P1 + k.iD12 * P2);
        o.f3[r] = -k.vol * k.iD22 * P2;
        o.f1[r] = -(o.f2[r] + o.f3[r]);
    }
}

// return mappings + stress for traditional particles (mpm_utils.py:1047-1103, 212-399, 8-84)
constexpr int STRESS_T_WB = (TF_F + S_F) * 32 * 4;
constexpr int STRESS_T_NW = 4;
__global__ void __launch_bounds__(32 * STRESS_T_NW) k_stress_traditional(int Nt, float* __restrict__ TF, float* __restrict__ TS,
                                                                         ModelDev md, float dt) {
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<STRESS_T_NW, STRESS_T_WB>(w, Nt, smem)) return;
    float* sT = reinterpret_cast<float*>(w.buf);
    float* sS = sT + 32 * TF_F;
    slab_load(w, {Slab{sT, TF, TF_F}});
    if (w.lane < w.cnt) {
        float* a = sT + w.lane * TF_F;
        float F[9], Ft[9], U[9], V[9], sg[3];
        float mu = a[T_MU], lam = a[T_LAM], ys = a[T_YS];
        const int mat = md.material;
#pragma unroll
        for (int i = 0; i < 9; i++) { Ft[i] = a[T_FT + i]; F[i] = Ft[i]; }
        if (mat == 1 || mat == 5) {  // von Mises (:212-255) / with damage (:258-311)
            svd3(Ft, U, sg, V);
            float sc[3], eps[3];
#pragma unroll
            for (int i = 0; i < 3; i++) { sc[i] = fmaxf(sg[i], 0.01f); eps[i] = logf(sc[i]); }
            float tr = eps[0] + eps[1] + eps[2], temp = tr / 3.0f;
            float tau[3];
#pragma unroll
            for (int i = 0; i < 3; i++) tau[i] = 2.0f * mu * eps[i] + lam * tr;
            float st = tau[0] + tau[1] + tau[2];
            float cn = len3(tau[0] - st / 3.0f, tau[1] - st / 3.0f, tau[2] - st / 3.0f);
            if (cn > ys && !(mat == 5 && ys <= 0.0f)) {
                float eh[3] = {eps[0] - temp, eps[1] - temp, eps[2] - temp};
                float ehn = len3(eh[0], eh[1], eh[2]) + 1e-6f;
                float dg = ehn - ys / (2.0f * mu);
                float corr[3] = {(dg / ehn) * eh[0], (dg / ehn) * eh[1], (dg / ehn) * eh[2]};
                float se[3];
#pragma unroll
                for (int i = 0; i < 3; i++) se[i] = expf(eps[i] - corr[i]);
                if (mat == 5) {
                    ys = ys - md.softening * len3(corr[0], corr[1], corr[2]);
                    if (ys <= 0.0f) { mu = 0.0f; lam = 0.0f; }
                }
                diag_sandwich(U, se, V, F);
                if (md.hardening == 1) ys = ys + 2.0f * mu * md.xi * dg;
            }
        } else if (mat == 2) {  // sand (:362-399)
            svd3(Ft, U, sg, V);
            float eps[3];
#pragma unroll
            for (int i = 0; i < 3; i++) eps[i] = logf(fmaxf(fabsf(sg[i]), 1e-14f));
            float tr = eps[0] + eps[1] + eps[2];
            float eh[3] = {eps[0] - tr / 3.0f, eps[1] - tr / 3.0f, eps[2] - tr / 3.0f};
            float ehn = len3(eh[0], eh[1], eh[2]);
            float dg = ehn + (3.0f * lam + 2.0f * mu) / (2.0f * mu) * tr * md.alpha;
            if (dg > 0.0f && tr > 0.0f) mat_mul_bt(U, V, F);
            if (dg > 0.0f && tr <= 0.0f) {
                float sn[3];
#pragma unroll
                for (int i = 0; i < 3; i++) sn[i] = expf(eps[i] - eh[i] * (dg / ehn));
                diag_sandwich(U, sn, V, F);
            }
        } else if (mat == 3) {  // viscoplastic StVK (:315-359)
            svd3(Ft, U, sg, V);
            float sc[3], eps[3], b[3];
#pragma unroll
            for (int i = 0; i < 3; i++) { sc[i] = fmaxf(sg[i], 0.01f); b[i] = sc[i] * sc[i]; eps[i] = logf(sc[i]); }
            float tr = eps[0] + eps[1] + eps[2];
            float st[3] = {2.0f * mu * (eps[0] - tr / 3.0f), 2.0f * mu * (eps[1] - tr / 3.0f), 2.0f * mu * (eps[2] - tr / 3.0f)};
            float stn = len3(st[0], st[1], st[2]);
            float y = stn - sqrtf(2.0f / 3.0f) * ys;
            if (y > 0.0f) {
                float mu_hat = mu * (b[0] + b[1] + b[2]) / 3.0f;
                float snn = stn - y / (1.0f + md.plastic_viscosity / (2.0f * mu_hat * dt));
                float se[3];
#pragma unroll
                for (int i = 0; i < 3; i++) se[i] = expf(1.0f / (2.0f * mu) * ((snn / stn) * st[i]) + tr / 3.0f);
                diag_sandwich(U, se, V, F);
            }
        }
        // stress from F (:1072-1103)
        float J = det3(F);
        svd3(F, U, sg, V);
        float S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (mat == 0 || mat == 5) {  // fixed corotated (:8-15)
            float Rm[9], D[9];
            mat_mul_bt(U, V, Rm);
#pragma unroll
            for (int i = 0; i < 9; i++) D[i] = F[i] - Rm[i];
            mat_mul_bt(D, F, S);
            float pj = lam * J * (J - 1.0f);
#pragma unroll
            for (int i = 0; i < 9; i++) S[i] *= 2.0f * mu;
            S[0] += pj; S[4] += pj; S[8] += pj;
        } else if (mat == 1 || mat == 3) {  // StVK / Hencky (:50-66)
            float sc[3], tau[3], t[9];
#pragma unroll
            for (int i = 0; i < 3; i++) sc[i] = fmaxf(sg[i], 0.01f);
            float sum = logf(sc[0]) + logf(sc[1]) + logf(sc[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) tau[i] = 2.0f * mu * logf(sc[i]) + lam * sum;
            diag_sandwich(U, tau, V, t);
            mat_mul_bt(t, F, S);
        } else if (mat == 2) {  // Drucker-Prager (:69-84)
            float sum = logf(sg[0]) + logf(sg[1]) + logf(sg[2]);
            float cc[3], t[9];
#pragma unroll
            for (int i = 0; i < 3; i++) cc[i] = 2.0f * mu * logf(sg[i]) * (1.0f / sg[i]) + lam * sum * (1.0f / sg[i]);
            diag_sandwich(U, cc, V, t);
            mat_mul_bt(t, F, S);
        }
        float* so = sS + w.lane * S_F;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) so[3 * r + cc] = (S[3 * r + cc] + S[3 * cc + r]) / 2.0f;
#pragma unroll
        for (int i = 0; i < 9; i++) a[T_F + i] = F[i];
        a[T_MU] = mu;
        a[T_LAM] = lam;
        a[T_YS] = ys;
    }
    slab_store(w, {Slab{sT, TF, TF_F}, Slab{sS, TS, S_F}});
}

// pre-P2G particle operations on one class (mpm_solver.py:260-279).  Position, velocity and mass of particle p sit at
// xr + p*xF, vr + p*vF and mr + p*mF (TP / VP records, or the element streams XE / EV / EFM.w).
__global__ void k_particle_ops(int n, const float* __restrict__ xr, int xF, float* __restrict__ vr, int vF, const float* __restrict__ mr,
                               int mF, const uint32_t* __restrict__ perm, int canon_offset, const ParticleOp* __restrict__ ops, int n_ops,
                               const StepState* __restrict__ st, float dt) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float time = (float)st->time;
    int ci = canon_offset + (int)perm[p];
    float* r = vr + (size_t)p * vF;
    const float* xp = xr + (size_t)p * xF;
    float vx = r[0], vy = r[1], vz = r[2];
    float m = mr[(size_t)p * mF];
    bool ch = false;
    for (int k = 0; k < n_ops; k++) {
        ParticleOp op = ops[k];
        if (!(time >= op.start_time && time < op.end_time)) continue;
        int mk = op.mask[ci];
        if (op.kind == 0 && mk == 1) { vx += op.vec[0] / m * dt; vy += op.vec[1] / m * dt; vz += op.vec[2] / m * dt; ch = true; }
        if (op.kind == 1 && mk >= 1) { vx += op.vec[0] * dt; vy += op.vec[1] * dt; vz += op.vec[2] * dt; ch = true; }
        if (op.kind == 2 && mk == 1) { vx = op.vec[0]; vy = op.vec[1]; vz = op.vec[2]; ch = true; }
        if (op.kind == 3 && mk == 1) {  // mpm_solver.py:1216-1254
            const float ox = xp[0] - op.point[0], oy = xp[1] - op.point[1], oz = xp[2] - op.point[2];
            const float dn = ox * op.n[0] + oy * op.n[1] + oz * op.n[2];
            const float hd = len3(ox - dn * op.n[0], oy - dn * op.n[1], oz - dn * op.n[2]);
            float theta = acosf((ox * op.h1[0] + oy * op.h1[1] + oz * op.h1[2]) / hd);
            if (!(ox * op.h2[0] + oy * op.h2[1] + oz * op.h2[2] > 0.0f)) theta = -theta;
            const float a1 = -hd * sinf(theta) * op.rot, a2 = hd * cosf(theta) * op.rot;
            vx = a1 * op.h1[0] + a2 * op.h2[0] + op.trans * op.n[0];
            vy = a1 * op.h1[1] + a2 * op.h2[1] + op.trans * op.n[1];
            vz = a1 * op.h1[2] + a2 * op.h2[2] + op.trans * op.n[2];
            ch = true;
        }
    }
    if (ch) { r[0] = vx; r[1] = vy; r[2] = vz; }
}

// ============================================================ P2G (+ fused cloth stress)
// p2g_apic_with_stress (mpm_utils.py:484-557), restructured for cell-sorted particles; for cloth
// elements compute_stress_from_F_trial (mpm_utils.py:1017-1046) is fused in front of it, so the
// stress never round-trips through HBM.
//  stage 1  lane = particle.  (elements: corner gather, return mapping + stress; the three corner forces leave as
//           REDG.128 into VF.)  The contribution of a
//           particle to stencil node (i,j,k), {dt*f + w m (v + C dpos), w m}, is separable:
//             out(i,j,k) = w2k T_ij + (w0i w1j) Uz_k,      T_ij = w1j Ux_i + w0i Uy_j,
//             Ux_i = w0i (A + i Bx) + dw0i S0,  Uy_j = w1j j By + dw1j S1,  Uz_k = w2k k Bz + dw2k S2
//           (A = {m (v - dx C f) + dt f_vertex, m}, B_a = m dx C[:,a], S_a = -dt/dx stress[:,a]); each lane
//           writes the 9 T_ij and the 3 {Uz_k / m, w2k} of ITS particle to shared memory (192 B per particle instead
//           of 27 float4; w0i w1j is recovered from the mass component of T_ij, see below).
//  stage 2  lane = stencil node (27 of 32 lanes): particles of one cell form a run; the lane expands and
//           accumulates its node along the run (2 broadcast LDS.128 + 3 FFMA2 + 1 FFMA per particle) and
//           flushes with ONE REDG.E.ADD.F32x4 per node per run; the node address is the run's base node plus a
//           per-lane constant (linear node layout).  No intra-warp reduction, no shared-memory
//           atomics; global atomics drop from 27*4 per particle to 27 per cell run.
#ifndef MPM_P2G_NW
#define MPM_P2G_NW 1
#endif
constexpr int P2G_NW = MPM_P2G_NW;  // warps (= slabs) per CTA
#ifndef MPM_P2G_V_NW
#define MPM_P2G_V_NW 1
#endif
#ifndef MPM_P2G_V_WARPS
#define MPM_P2G_V_WARPS 32
#endif
constexpr int P2G_V_NW = MPM_P2G_V_NW, P2G_V_MINB = MPM_P2G_V_WARPS / MPM_P2G_V_NW;  // the vertex scatter
// Stage 2 can take a run's node offsets from its leader lane (as the gather does) instead of decoding the packed cell in
// every lane: 20 instructions less per flush, but the 1152 bytes of offsets per warp cost resident warps (31 instead of 35
// by shared memory) -- measured on C3: 88.4 us per substep with, 86.8 without.
#ifndef MPM_P2G_LEADER_OFFSETS
#define MPM_P2G_LEADER_OFFSETS 0
#endif
// T tiles | U tiles | per-run node offsets (or the 32 cells of the slab in grouped order, which is also the slack the
// stage-2 look-ahead loads run into)
constexpr int P2G_T_B = 32 * 9 * 16, P2G_U_B = 32 * 3 * 16, P2G_O_B = MPM_P2G_LEADER_OFFSETS ? 32 * 9 * 4 : 128;
constexpr int P2G_WB = P2G_T_B + P2G_U_B + P2G_O_B;  // 7296 B; the raw slabs of the record kernels (<= 3328 B) are overlaid on the tiles
constexpr int P2G_SMEM = 128 + P2G_NW * P2G_WB;  // the stage-2 look-ahead reads a few bytes past the U tiles: into the offsets

struct P2GPart {  // what stage 1 hands to the tile writer, lane = particle
    float x[3], m, v[3], C[9];
    float Sp[9];  // -dt/dx * (vol *) stress, row-major; unused when STRESS = false
};

// weights + contribution tiles of this lane's particle, written in cell-grouped order; returns the packed stencil base
// cell of SLOT `lane` of that order (a distinct negative value for the trailing slots without a particle).  fetch_force(fv) is called after the weights are done: the vertex scatter
// waits for its predecessor there (the vertex forces are the only input the element kernel produces).
template <bool STRESS, typename FV>
__device__ __forceinline__ int p2g_write_tiles(const Grid& g, const Warp& w, bool valid, P2GPart& P, float dt, float rpic,
                                               FV&& fetch_force) {
    float* C = P.C;
    if (rpic != 0.0f) {  // mpm_utils.py:528-542
        float Cn[9];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
                Cn[3 * a + b] = (1.0f - rpic) * C[3 * a + b] + rpic / 2.0f * (C[3 * a + b] - C[3 * b + a]);
#pragma unroll
        for (int i = 0; i < 9; i++) C[i] = (rpic < -0.001f) ? 0.0f : Cn[i];
    }
    int b[3];
    float f[3], wgt[3][3], dwg[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float gp = P.x[a] * g.inv_dx;
        b[a] = (int)(gp - 0.5f);
        f[a] = gp - (float)b[a];
#pragma unroll
        for (int i = 0; i < 3; i++) bspline(f[a], i, wgt[a][i], dwg[a][i]);
    }
    float fv[3] = {0.f, 0.f, 0.f};
    fetch_force(fv);
    const float m = P.m;
    // w m (v + C dpos) + dt w f_vertex = w (A + B_x i + B_y j + B_z k), dpos = (ijk - f) dx
    const float mdx = m * g.dx;
    float A[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
        A[c] = m * (P.v[c] - g.dx * (C[3 * c] * f[0] + C[3 * c + 1] * f[1] + C[3 * c + 2] * f[2])) + dt * fv[c];
    const V4 A4 = v4(A[0], A[1], A[2], m);
    const V4 Bx = v4(mdx * C[0], mdx * C[3], mdx * C[6], 0.f), By = v4(mdx * C[1], mdx * C[4], mdx * C[7], 0.f),
             Bz = v4(mdx * C[2], mdx * C[5], mdx * C[8], 0.f);
    V4 Ux[3], Uy[3], Uz[3];
    Ux[0] = mul4(wgt[0][0], A4);
    Ux[1] = mul4(wgt[0][1], add4(A4, Bx));
    Ux[2] = mul4(wgt[0][2], fma4(2.0f, Bx, A4));
    Uy[0] = v4(0.f, 0.f, 0.f, 0.f);
    Uy[1] = mul4(wgt[1][1], By);
    Uy[2] = mul4(2.0f * wgt[1][2], By);
    Uz[0] = v4(0.f, 0.f, 0.f, 0.f);
    Uz[1] = mul4(wgt[2][1], Bz);
    Uz[2] = mul4(2.0f * wgt[2][2], Bz);
    if (STRESS) {  // dt * (-stress grad w), stress pre-scaled in Sp
        const float* Sp = P.Sp;
        const V4 S0 = v4(Sp[0], Sp[3], Sp[6], 0.f), S1 = v4(Sp[1], Sp[4], Sp[7], 0.f), S2 = v4(Sp[2], Sp[5], Sp[8], 0.f);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Ux[i] = fma4(dwg[0][i], S0, Ux[i]);
            Uy[i] = fma4(dwg[1][i], S1, Uy[i]);
            Uz[i] = fma4(dwg[2][i], S2, Uz[i]);
        }
    }
    // The second term of out(i,j,k) needs w0i w1j, and the mass component of T_ij IS m w0i w1j: with Uz_k stored
    // as Uz_k / m the term becomes T_ij.w * (Uz_k / m) and stage 2 reads 8 words per particle instead of 9 (no
    // separate weight tile).  m = 0 (ghost vertices of a sharded run): Uz_k = 0 for vertices, exact.  A massless
    // element / traditional particle (volume > 0, density 0) still exerts its stress force: rare slow path below.
    const float inv_m = (m != 0.0f) ? 1.0f / m : 0.0f;
    if (STRESS && valid && m == 0.0f) {
#pragma unroll  // static indices: a rolled loop would push wgt / Uz into local memory for the whole kernel
        for (int n = 0; n < 27; n++) {
            const int i = n / 9, j = (n / 3) % 3, k = n % 3;
            const float wij = wgt[0][i] * wgt[1][j];
            const int ni = node_index(g, b[0] + i, b[1] + j, b[2] + k);
            if (ni >= 0) atomicAdd(&g.acc[ni], make_float4(wij * Uz[k].lo.x, wij * Uz[k].lo.y, wij * Uz[k].hi.x, 0.f));
        }
    }
    // the slab is re-grouped by CURRENT cell on the way into the tiles: this particle's record goes to slot G.slot, so
    // that stage 2 meets every distinct cell as ONE contiguous run however far the particles have drifted since the sort
    const int mycell = valid ? pack_cell(b[0], b[1], b[2]) : -1 - w.lane;
    const Runs G = group_runs(w.lane, w.cnt, mycell);
    float4* tT = reinterpret_cast<float4*>(w.buf) + G.slot * 9;
    float4* tU = reinterpret_cast<float4*>(w.buf + P2G_T_B) + G.slot * 3;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const V4 T = fma4(wgt[1][j], Ux[i], mul4(wgt[0][i], Uy[j]));
            tT[i * 3 + j] = make_float4(T.lo.x, T.lo.y, T.hi.x, T.hi.y);
        }
#pragma unroll
    for (int k = 0; k < 3; k++) tU[k] = make_float4(Uz[k].lo.x * inv_m, Uz[k].lo.y * inv_m, Uz[k].hi.x * inv_m, wgt[2][k]);
    // particles whose stencil leaves the grid are flagged (the reference has no bounds check there, mpm_utils.py:516-557)
    const bool inside = (unsigned)b[0] <= (unsigned)(g.n - 3) && (unsigned)b[1] <= (unsigned)(g.n - 3) && (unsigned)b[2] <= (unsigned)(g.n - 3);
    if (valid && !inside) g.flags[1] = 1;
    // the cell of every SLOT, for stage 2
    int* cs = reinterpret_cast<int*>(w.buf + P2G_T_B + P2G_U_B);
    cs[G.slot] = mycell;
    __syncwarp();
    return cs[w.lane];
}

// stage 2: lane = stencil node, one pass over the slab's 32 tile records, one REDG.128 per node per cell run
__device__ __forceinline__ void p2g_stage2(const Grid& g, const Warp& w, int mycell) {
    const Runs R = find_runs(w.lane, w.cnt, mycell);
#if MPM_P2G_LEADER_OFFSETS
    int* ro = reinterpret_cast<int*>(w.buf + P2G_T_B + P2G_U_B);
    __syncwarp();  // the slot cells in this region have been read
    if ((R.starts >> w.lane) & 1u) stencil_offsets(g, mycell, ro + R.mine * 9);  // the run's leader: nine per-axis node offsets
#endif
    __syncwarp();
    const bool act = w.lane < 27;
    const LaneNode ln(w.lane);  // lanes 27..31 shadow lane 26 (same shared-memory words: no extra wavefront)
    const float4* pT = reinterpret_cast<const float4*>(w.buf) + (ln.li * 3 + (ln.lj - 3));
    const float4* pU = reinterpret_cast<const float4*>(w.buf + P2G_T_B) + (ln.lk - 6);
    int run = 0;
    int c = __shfl_sync(0xffffffffu, mycell, 0);  // packed cell of the current run (used without leader offsets)
    auto flush = [&](float2 lo, float2 hi) {
        if (act && (hi.y != 0.0f || lo.x != 0.0f || lo.y != 0.0f || hi.x != 0.0f)) {
#if MPM_P2G_LEADER_OFFSETS
            bool ok;
            const int ni = ln.node(ro + run * 9, ok);
#else  // every lane decodes the run's packed cell itself
            const int ni = node_index(g, (c & 1023) - 2 + ln.li, ((c >> 10) & 1023) - 2 + (ln.lj - 3), ((c >> 20) & 1023) - 2 + (ln.lk - 6));
            const bool ok = ni >= 0;
#endif
#ifdef MPM_NO_RED  // analysis build: everything but the atomics (results are wrong)
            if (ni == 0x7fffffff) atomicAdd(&g.acc[0], make_float4(lo.x, lo.y, hi.x, hi.y));
#else
            if (ok) atomicAdd(&g.acc[ni], make_float4(lo.x, lo.y, hi.x, hi.y));
#endif
        }
        run++;
    };
    float2 alo = make_float2(0.f, 0.f), ahi = alo;
    // one pass over the slab with a fixed trip count; the shared-memory loads of particle q+1 are in flight while
    // particle q is accumulated (deeper prefetch was measured: no gain, more address arithmetic).  The loads of
    // "particle 32" read a few bytes past the tiles, inside the warp's region, and are never used.  A warp-uniform
    // test of `starts` closes a run.
    float4 Tn = pT[0], Un = pU[0];
#pragma unroll 8
    for (int q = 0; q < 32; q++) {  // records past cnt contribute zeros
        const float4 T = Tn, U = Un;
        Tn = pT[(q + 1) * 9];
        Un = pU[(q + 1) * 3];
        if (q > 0 && ((R.starts >> q) & 1u)) {
            flush(alo, ahi);
            alo = ahi = make_float2(0.f, 0.f);
#if !MPM_P2G_LEADER_OFFSETS
            c = __shfl_sync(0xffffffffu, mycell, q);
#endif
        }
        alo = fma2(U.w, make_float2(T.x, T.y), alo);
        ahi = fma2(U.w, make_float2(T.z, T.w), ahi);
        alo = fma2(T.w, make_float2(U.x, U.y), alo);  // (m w0i w1j) (Uz_k / m)
        ahi.x = fmaf(T.w, U.z, ahi.x);
    }
    flush(alo, ahi);
}

struct P2GIn {
    const float* KP;  // TP / VP records
    const float* SF;  // TS (KIND 1) or the vertex forces VF (KIND 2)
};

// ---- cloth elements: constitutive update + scatter in one kernel.  The element state arrives as coalesced float4
// streams (lane = particle, one LDG.128 each; see mpm_device.cuh) -- no shared-memory staging, the tiles of stage 1
// are the only shared-memory traffic.
struct ElemIO {
    const int4* EFM;
    const float4 *K0, *K1;
    const float4 *XE, *EV, *ED1, *ED2, *C0, *C1;
    float4* D3;   // buffer `cur`: read, return-mapped in place
    float4* SP3;
    float4* VF;   // buffer `cur`
};
#ifndef MPM_P2G_E_WARPS
#define MPM_P2G_E_WARPS 28
#endif
__global__ void __launch_bounds__(32 * P2G_NW, MPM_P2G_E_WARPS / P2G_NW) k_p2g_elements(Grid g, ElemIO A, int n, float dt, float rpic, float friction_coeff) {
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<P2G_NW, P2G_WB>(w, n, smem)) return;
    ts_begin(g, TS_P2G_E);
    PHASE_BEGIN();
    const bool valid = w.lane < w.cnt;
    const int p = w.p0 + w.lane;  // the arrays carry 32 records of slack: lanes past cnt read (and discard) in bounds
    // constants: nobody writes them between imports
    const int4 efm = __ldg(&A.EFM[p]);
    const float4 k0 = __ldg(&A.K0[p]), k1 = __ldg(&A.K1[p]);
    pdl_wait();     // the element G2P of the previous substep (the last kernel of a substep) is complete
    pdl_trigger();
    const float4 xe = A.XE[p], ev = A.EV[p], e1 = A.ED1[p], e2 = A.ED2[p], c0 = A.C0[p], c1 = A.C1[p], d3v = A.D3[p];
    P2GPart P;
    P.x[0] = xe.x; P.x[1] = xe.y; P.x[2] = xe.z;
    P.v[0] = ev.x; P.v[1] = ev.y; P.v[2] = ev.z;
    P.m = __int_as_float(efm.w);
    P.C[0] = c0.x; P.C[1] = c0.y; P.C[2] = c0.z; P.C[3] = c0.w; P.C[4] = c1.x; P.C[5] = c1.y; P.C[6] = c1.z; P.C[7] = c1.w; P.C[8] = d3v.w;
    if (!valid) {  // a harmless stand-in: zero contribution, finite arithmetic
        P.m = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) { P.x[i] = 0.f; P.v[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < 9; i++) { P.C[i] = 0.f; P.Sp[i] = 0.f; }
    }
    PHASE(g, 0, 0);  // loads
    if (valid) {
        const float d1[3] = {e1.x, e1.y, e1.z}, d2[3] = {e2.x, e2.y, e2.z}, d3[3] = {d3v.x, d3v.y, d3v.z};
        const ElemConst ek{k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        ElemStress es;
        element_stress(d1, d2, d3, ek, friction_coeff, es);
        // vertex_force scatter (mpm_utils.py:172-175): one 16-byte vector atomic per corner
        atomicAdd(&A.VF[efm.x], make_float4(es.f1[0], es.f1[1], es.f1[2], 0.f));
        atomicAdd(&A.VF[efm.y], make_float4(es.f2[0], es.f2[1], es.f2[2], 0.f));
        atomicAdd(&A.VF[efm.z], make_float4(es.f3[0], es.f3[1], es.f3[2], 0.f));
        A.D3[p] = make_float4(es.nd3[0], es.nd3[1], es.nd3[2], d3v.w);
        // stress = vol * P3 (x) nd3 already carries the volume (mpm_utils.py:177, :494); kept for state.particle_stress
        A.SP3[p] = make_float4(ek.vol * es.P3[0], ek.vol * es.P3[1], ek.vol * es.P3[2], 0.f);
        const float sc = -dt * g.inv_dx * ek.vol;
#pragma unroll
        for (int rr = 0; rr < 3; rr++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) P.Sp[3 * rr + cc] = sc * (es.P3[rr] * es.nd3[cc]);
    }
    PHASE(g, 0, 1);  // stress
    const int mycell = p2g_write_tiles<true>(g, w, valid, P, dt, rpic, [](float (&)[3]) {});
    PHASE(g, 0, 2);  // stage 1
    p2g_stage2(g, w, mycell);
    PHASE(g, 0, 3);  // stage 2
    PHASE_END(g, 0);
    ts_end(g, TS_P2G_E);
}

// ============================================================ collider / mover scatter
struct Stencil {
    int b[3];
    float w[3][3];
};
__device__ __forceinline__ void make_stencil(const Grid& g, float x, float y, float z, Stencil& s) {
    float p[3] = {x * g.inv_dx, y * g.inv_dx, z * g.inv_dx};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        s.b[a] = (int)(p[a] - 0.5f);
        float f = p[a] - (float)s.b[a];
        float wa = 1.5f - f, wb = f - 1.0f, wc = f - 0.5f;
        s.w[a][0] = wa * wa * 0.5f;
        s.w[a][1] = 0.0f - wb * wb + 0.75f;
        s.w[a][2] = wc * wc * 0.5f;
    }
}
// bounds test of compute_mesh / add_velocity_* (mpm_solver.py:692,858)
__device__ __forceinline__ bool scatter_ok(const Grid& g, const Stencil& s) {
    return s.b[0] >= 0 && s.b[0] < g.n - 3 && s.b[1] >= 0 && s.b[1] < g.n - 3 && s.b[2] >= 0 && s.b[2] < g.n - 3;
}

// compute_mesh (mpm_solver.py:829-880).  Only nodes of ALLOCATED blocks are written: a node
// outside every particle stencil is never read by G2P, so the body mesh never allocates grid
// (and faces far from the cloth cost eight table reads and nothing else).
struct ColliderArgs {
    int Mf;
    const int* faces;
    const float *px, *pv;
    const StepState* st;
    float dt;
    int advance;
};
// one scatter of a 3^3 stencil into the ACTIVE blocks only: node = X(ix) + Y(iy) + Z(iz) (axis_offset); the activity of
// the (up to) eight blocks under the stencil is an 8-bit mask, a node's block is picked by three per-axis crossing bits
// (i_only >= 0: only the nodes of stencil plane i = i_only -- the scatter kernels give every face / joint particle three
// threads, one per plane, which shortens the chain of dependent atomics of a thread from 27 (x2) to 9 (x2))
template <typename F>
__device__ __forceinline__ void scatter_active_nodes(const Grid& g, const Stencil& sp, unsigned act8, bool flag_inactive, int i_only, F&& add) {
    int ox[3], oy[3], oz[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        ox[i] = axis_offset<0>(g, sp.b[0] + i);
        oy[i] = axis_offset<1>(g, sp.b[1] + i);
        oz[i] = axis_offset<2>(g, sp.b[2] + i);
    }
    const unsigned cx = axis_cross(sp.b[0]), cy = axis_cross(sp.b[1]), cz = axis_cross(sp.b[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i_only >= 0 && i != i_only) continue;
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const unsigned c = (((cx >> i) & 1u) << 2) | (((cy >> j) & 1u) << 1) | ((cz >> k) & 1u);
                if (!((act8 >> c) & 1u)) {
                    if (flag_inactive) g.flags[1] = 1;
                    continue;
                }
                add(ox[i] + oy[j] + oz[k], sp.w[0][i] * sp.w[1][j] * sp.w[2][k]);
            }
    }
}
__device__ __forceinline__ void collider_scatter_face(const Grid& g, const ColliderArgs& ca, int f, int i_only = -1) {
    const int* __restrict__ faces = ca.faces;
    const float* __restrict__ px = ca.px;
    const float* __restrict__ pv = ca.pv;
    const StepState* __restrict__ st = ca.st;
    const float dt = ca.dt;
    const int advance = ca.advance;
    const float s = advance ? (float)((double)dt * (double)st->k) : 0.0f;
    int id[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    float P[3][3], fv[3] = {0, 0, 0}, fp[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float vel = pv[3 * id[c] + a];
            P[c][a] = px[3 * id[c] + a] + s * vel;
            fv[a] += vel;
            fp[a] += P[c][a];
        }
#pragma unroll
    for (int a = 0; a < 3; a++) { fv[a] = fv[a] / 3.0f; fp[a] = fp[a] / 3.0f; }
    Stencil sp;
    make_stencil(g, fp[0], fp[1], fp[2], sp);
    if (!scatter_ok(g, sp)) return;
    const unsigned act8 = load_active8(g, sp.b[0], sp.b[1], sp.b[2]);
    if (!act8) return;
    float e1[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
    float e2[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
    float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    float nl = len3(nx, ny, nz);
    if (nl > 0.0f) { nx /= nl; ny /= nl; nz /= nl; } else { nx = ny = nz = 0.0f; }
    scatter_active_nodes(g, sp, act8, false, i_only, [&](int ni, float ww) {
        atomicAdd(&g.colv[ni], make_float4(ww * fv[0], ww * fv[1], ww * fv[2], ww));
        atomicAdd(&g.coln[ni], make_float4(ww * nx, ww * ny, ww * nz, 0.0f));
    });
}
__global__ void __launch_bounds__(128) k_collider_scatter(Grid g, ColliderArgs ca) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < ca.Mf) collider_scatter_face(g, ca, f);
}

// add_velocity_traditional / _verts / _faces (mpm_solver.py:677-788) in one launch:
// threads [0,njt) pinned traditional tail, [njt, njt+njv) joint vertices, then joint faces.
struct MoverArgs {
    int njt, njv, njf, Nt;
    const float *vt, *vvv, *vf, *TP, *VP;
    const float4* XE;
    const int *invE, *invT, *invV;
};
__device__ __forceinline__ void mover_scatter_one(const Grid& g, const MoverArgs& ma, int t, int i_only = -1) {
    const int njt = ma.njt, njv = ma.njv, Nt = ma.Nt;
    const float *__restrict__ vt = ma.vt, *__restrict__ vvv = ma.vvv, *__restrict__ vf = ma.vf;
    const float *__restrict__ TP = ma.TP, *__restrict__ VP = ma.VP;
    const int *__restrict__ invE = ma.invE, *__restrict__ invT = ma.invT, *__restrict__ invV = ma.invV;
    const float* xr;
    const float* vel;
    if (t < njt) { xr = TP + (size_t)invT[Nt - njt + t] * KP_F; vel = vt + 3 * t; }
    else if (t < njt + njv) { xr = VP + (size_t)invV[t - njt] * VP_F; vel = vvv + 3 * (t - njt); }
    else { xr = reinterpret_cast<const float*>(ma.XE + invE[t - njt - njv]); vel = vf + 3 * (t - njt - njv); }
    Stencil sp;
    make_stencil(g, xr[0], xr[1], xr[2], sp);
    if (!scatter_ok(g, sp)) return;
    const unsigned act8 = load_active8(g, sp.b[0], sp.b[1], sp.b[2]);
    const float v0 = vel[0], v1 = vel[1], v2 = vel[2];
    scatter_active_nodes(g, sp, act8, true, i_only, [&](int ni, float ww) {
        atomicAdd(&g.mov[ni], make_float4(ww * v0, ww * v1, ww * v2, ww));
    });
}
__global__ void __launch_bounds__(128) k_mover_scatter(Grid g, MoverArgs ma) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ma.njt + ma.njv + ma.njf) mover_scatter_one(g, ma, t);
}
// Both scatters in one launch, placed behind the vertex P2G in the stream.  It depends on neither P2G kernel
// (different accumulators; the block table and the positions are those of the previous substep), so its CTAs fill
// the tail of the vertex P2G; the wait before exit keeps completion transitive.  (Placing it between the two P2G
// kernels was measured: its ~1.1 M vector atomics then collide with the element kernel's and slow that kernel by
// more than the scatter's own exposed time.)  Thread triples [0, Mf) faces, then the movers.
__device__ __forceinline__ void body_scatter_thread(const Grid& g, const ColliderArgs& ca, const MoverArgs& ma, int tt) {
    const int t = tt / 3, i_only = tt - 3 * t;  // three threads per face / joint particle: one stencil plane each
    if (t < ca.Mf) collider_scatter_face(g, ca, t, i_only);
    else if (t - ca.Mf < ma.njt + ma.njv + ma.njf) mover_scatter_one(g, ma, t - ca.Mf, i_only);
}
__global__ void __launch_bounds__(128) k_body_scatter(Grid g, ColliderArgs ca, MoverArgs ma) {
    pdl_trigger();
    ts_begin(g, TS_SCATTER);
    body_scatter_thread(g, ca, ma, blockIdx.x * blockDim.x + threadIdx.x);
    ts_end(g, TS_SCATTER);  // before the wait: the stamp is the end of this kernel's own work
    pdl_wait();
}
// The same work as CTAs [first, first + n_ctas) of the vertex P2G grid (k_p2g<2>): placed in the MIDDLE of that grid they
// start after the element kernel has drained (no collision of their ~1.1 M vector atomics with its own) and finish before
// the vertex kernel does, so the scatter leaves the critical path and one launch disappears.
struct BodyScatter {
    ColliderArgs ca;
    MoverArgs ma;
    int first, n_ctas;
};

// KIND 1: traditional (stress*vol, :496), 2: cloth vertex.  PDL invariants of this file: (1) every kernel executes
// griddepcontrol.wait before it exits, so completion is transitive along the stream; (2) a kernel whose successor
// runs code in front of its own wait that READS what this kernel's predecessors wrote triggers only AFTER its
// own wait, so "my grid has started" implies "my predecessor's predecessor has completed".
template <int KIND>
__global__ void __launch_bounds__(32 * (KIND == 2 ? P2G_V_NW : P2G_NW), KIND == 2 ? P2G_V_MINB : 1) k_p2g(Grid g, P2GIn in, int n, float dt, float rpic, BodyScatter bs) {
    static_assert(KIND == 1 || KIND == 2, "elements have their own kernel");
    constexpr int F0 = (KIND == 2) ? VP_F : KP_F;  // kinematics record
    constexpr int NW = (KIND == 2) ? P2G_V_NW : P2G_NW;
    int blk = blockIdx.x;
    if (KIND == 2 && bs.n_ctas > 0) {  // the body scatter rides in this grid: CTAs [first, first + n_ctas)
        const int rel = blk - bs.first;
        if (rel >= 0 && rel < bs.n_ctas) {
            // it depends on neither P2G kernel (different accumulators; the block table and the positions are those of the
            // previous substep, complete since this grid could only start after the element kernel's wait): no wait here,
            // the slab CTAs of this grid keep completion transitive
            if (g.ts && threadIdx.x == 0 && rel < 256) atomicMin(&g.ts[2 * TS_SCATTER], globaltimer_ns());
            body_scatter_thread(g, bs.ca, bs.ma, rel * (32 * NW) + (int)threadIdx.x);
            if (g.ts && threadIdx.x == 0 && rel + 256 >= bs.n_ctas) atomicMax(&g.ts[2 * TS_SCATTER + 1], globaltimer_ns());
            return;
        }
        if (rel >= bs.n_ctas) blk -= bs.n_ctas;
    }
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<NW, P2G_WB>(w, n, smem, blk)) return;
    ts_begin(g, TS_P2G_E + KIND);
    PHASE_BEGIN();
    float* buf = reinterpret_cast<float*>(w.buf);
    // The vertex scatter only needs its predecessor (the element kernel) for the vertex forces: the VP slab is
    // loaded and unpacked while the element kernel drains, griddepcontrol.wait sits in front of the VF read.
    if (KIND != 2) {
        pdl_wait();
        pdl_trigger();  // let the successor's CTAs be scheduled into the slots this grid frees
    }
    float* raw1 = buf + 32 * F0;
    if (KIND == 1) slab_load(w, {Slab{buf, in.KP, KP_F}, Slab{raw1, in.SF, S_F}});
    if (KIND == 2) slab_load(w, {Slab{buf, in.KP, VP_F}});
    PHASE(g, KIND, 0);  // slab load
    const bool valid = w.lane < w.cnt;
    P2GPart P;
    P.m = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) { P.x[i] = 0.f; P.v[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 9; i++) { P.C[i] = 0.f; P.Sp[i] = 0.f; }
    if (valid) {
        const float* r = buf + w.lane * F0;
        if (KIND == 2) {
            const float4* r4 = reinterpret_cast<const float4*>(r);
            const float4 a = r4[0], b = r4[1], c = r4[2], d = r4[3];
            P.x[0] = a.x; P.x[1] = a.y; P.x[2] = a.z; P.m = a.w;
            P.v[0] = b.x; P.v[1] = b.y; P.v[2] = b.z;
            P.C[0] = b.w; P.C[1] = c.x; P.C[2] = c.y; P.C[3] = c.z; P.C[4] = c.w; P.C[5] = d.x; P.C[6] = d.y; P.C[7] = d.z; P.C[8] = d.w;
        } else {
            P.x[0] = r[0]; P.x[1] = r[1]; P.x[2] = r[2]; P.m = r[P_M];
            P.v[0] = r[P_V]; P.v[1] = r[P_V + 1]; P.v[2] = r[P_V + 2];
#pragma unroll
            for (int i = 0; i < 9; i++) P.C[i] = r[P_C + i];
            const float sc = -dt * g.inv_dx * r[P_VOL];
            const float* s = raw1 + w.lane * S_F;
#pragma unroll
            for (int i = 0; i < 9; i++) P.Sp[i] = sc * s[i];
        }
    }
    __syncwarp();  // the contributions overwrite the raw slabs
    PHASE(g, KIND, 1);  // unpack
    const int mycell = p2g_write_tiles<KIND == 1>(g, w, valid, P, dt, rpic, [&](float (&fv)[3]) {
        if (KIND == 2) {  // the weights were computed while the element kernel drained
            pdl_wait();
            pdl_trigger();
            if (valid) {
                const float4 f4 = __ldcg(reinterpret_cast<const float4*>(in.SF) + w.p0 + w.lane);  // written by L2 atomics
                fv[0] = f4.x; fv[1] = f4.y; fv[2] = f4.z;
            }
        }
    });
    PHASE(g, KIND, 2);  // stage 1
    p2g_stage2(g, w, mycell);
    PHASE(g, KIND, 3);  // stage 2
    PHASE_END(g, KIND);
    ts_end(g, TS_P2G_E + KIND);
}


// ============================================================ grid update
// One pass over the nodes of the allocated blocks that fuses
//   grid_normalization_and_gravity (mpm_utils.py:561-572), add_damping_via_grid (:1162-1174),
//   mesh collider normalize_grid + collide (mpm_solver.py:882-917),
//   particle mover normalize_grid (:790-799), and every grid_postprocess BC in order (:487-501),
// then re-zeroes the accumulators it consumed (replaces the three dense zero_grid sweeps).
// PEER (sharded runs with the peer-to-peer exchange): the WHOLE exchange of the blocks shared with other ranks is fused in,
// as a flagged-data ("LL") protocol -- no fence, no counter, no separate kernel:
//   1. push: the threads sweep the shared-block lists; for every block this rank is a member of they take its partial sums
//      (acc, mov) out of the grid (re-zeroing them) and store them, each 8 bytes as {float, epoch flag}, straight into the
//      receive areas of ALL members of the block, this rank's own included (PeerArea in mpm_device.cuh);
//   2. the nodes of blocks that are not shared are updated while those stores are in flight;
//   3. the nodes of shared blocks are updated from the sum of the members' parts IN RANK ORDER (every member computes
//      identical bits); each part is polled until both of its flags carry this epoch, so a rank waits for exactly the
//      neighbours it shares the node with, not for a global barrier.
// Step 3 waits on CTAs of the peers' grid updates, so every CTA of this grid must be resident at once: the host launches
// 4 x 148 CTAs and the launch bounds pin 4 CTAs per SM.  Two epoch parities of receive area suffice: a rank can only push
// epoch e+2 after its grid update of e+1, which needed its neighbours' pushes of e+1, which follow their grid updates of e.
template <bool PEER>
__global__ void __launch_bounds__(256, PEER ? 4 : 5) k_grid_update(Grid g, ModelDev md, float dt, int use_collider, float col_friction,
                                                     int use_mover, const BCDesc* __restrict__ bcs, int n_bc,
                                                     const StepState* __restrict__ st, PeerArea P) {
    // the active list was last changed by the previous substep's G2P: the node address is computed while the
    // scatter kernels in front of this one drain (PDL invariant: wait, then trigger)
    ts_begin(g, TS_GRID);
    const int n_slots = min(*g.n_slots, g.cap);
    const int total = n_slots * BN;
    const int idx0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    int co_next = idx0 < total ? g.slot_coord[idx0 >> 6] : 0;
    unsigned E = 0;
    size_t par_off = 0;
    int n_push_a = 0, n_push = 0;
    if (PEER) {
        E = *P.epoch;  // advanced by the previous grid update, which completed long before this one was launched
        par_off = (size_t)(E & 1) * P.nranks * P.slot_bytes;
        n_push_a = min(*P.nA, P.capA) * BN;
        n_push = n_push_a + min(*P.nM, P.capM) * BN;
    }
    pdl_wait();     // the scatters of this substep
    pdl_trigger();  // the G2P kernels behind this one may take the slots it frees (they wait for its completion)
    if (PEER) {
        ts_begin(g, TS_PUSH);
        const unsigned flag = E + 1;
        for (int i = idx0; i < n_push; i += stride) {
            const bool m = i >= n_push_a;
            const int e = m ? i - n_push_a : i, j = e >> 6, l = e & 63;
            const int mem = (m ? P.memM : P.memA)[j];
            if (!((mem >> P.rank) & 1)) continue;  // not a member: this rank cannot reach the block
            const int co = (m ? P.listM : P.listA)[j];
            float4* src = (m ? g.mov : g.acc) + block_node(g, co, l);
            const float4 v = *src;  // zero if the block is not active here (invariant 2)
            if (v.w != 0.0f || v.x != 0.0f || v.y != 0.0f || v.z != 0.0f) *src = make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t off = par_off + (size_t)P.rank * P.slot_bytes + ((((size_t)(m ? P.capA + j : j)) * BN + l) << 5);
            for (int r = 0; r < P.nranks; r++)
                if ((mem >> r) & 1) ll_store(P.base[r] + off, v, flag);
        }
        ts_end(g, TS_PUSH);
    }
    const float time = (float)st->time;
    for (int pass = 0; pass < (PEER ? 2 : 1); pass++) {
    if (PEER && pass == 1) {
        ts_begin(g, TS_PULL);  // slot 9 of the sharded timeline: the shared nodes, incl. the wait for their members' parts
        co_next = idx0 < total ? g.slot_coord[idx0 >> 6] : 0;
    }
    for (int idx = idx0; idx < total; idx += stride) {
        const int l = idx & 63;
        const int co = co_next;
        {
            const int nidx = idx + stride;
            if (nidx < total) co_next = g.slot_coord[nidx >> 6];
        }
        const int ni = block_node(g, co, l);
        int ja = -1, jm = -1, memA = 0, memM = 0;
        if (PEER) {
            const int blk = ni >> 6;
            ja = P.mapA[blk]; jm = P.mapM[blk];
            if (ja >= 0) { memA = P.memA[ja]; if (!((memA >> P.rank) & 1)) ja = -1; }
            if (jm >= 0) { memM = P.memM[jm]; if (!((memM >> P.rank) & 1)) jm = -1; }
            if ((ja >= 0 || jm >= 0) != (pass == 1)) continue;
        }
        // all four accumulator loads are issued before any use (memory-level parallelism).  A quantity of a shared block was
        // taken out of the grid by the push: it comes from the receive area instead
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), mv = a;
        if (ja < 0) a = g.acc[ni];
        if (jm < 0) mv = g.mov[ni];
        const float4 a_own = a, mv_own = mv;  // what is still in the grid and has to be re-zeroed
        if (PEER) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int j = h ? jm : ja;
                if (j < 0) continue;
                const int mem = h ? memM : memA;
                const unsigned char* src = P.base[P.rank] + par_off + ((((size_t)(h ? P.capA + j : j)) * BN + l) << 5);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < P.nranks; r++) {
                    if (!((mem >> r) & 1)) continue;
                    const float4 v = ll_load(src + (size_t)r * P.slot_bytes, E + 1);
                    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                }
                if (h) mv = sum; else a = sum;
            }
        }
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), cn = cv;
        if (use_collider) { cv = g.colv[ni]; cn = g.coln[ni]; }
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (g.dbg_acc) g.dbg_acc[ni] = a;
        if (a.w > 1e-15f) {
            float inv = 1.0f / a.w;
            vx = a.x * inv + dt * md.gx;
            vy = a.y * inv + dt * md.gy;
            vz = a.z * inv + dt * md.gz;
        }
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_own.w != 0.0f || a_own.x != 0.0f || a_own.y != 0.0f || a_own.z != 0.0f) g.acc[ni] = zero4;
        if (md.damping < 1.0f) {
            vx -= (1.0f - md.damping) * vx;
            vy -= (1.0f - md.damping) * vy;
            vz -= (1.0f - md.damping) * vz;
        }
        if (use_collider && cv.w != 0.0f) {
            g.colv[ni] = zero4;
            g.coln[ni] = zero4;
            if (cv.w > 1e-15f) {
                float inv = 1.0f / cv.w;
                float mx = cv.x * inv, my = cv.y * inv, mz = cv.z * inv;
                float rx = vx - mx, ry = vy - my, rz = vz - mz;
                float nl = len3(cn.x, cn.y, cn.z);
                float nx = 0.f, ny = 0.f, nz = 0.f;
                if (nl > 0.0f) { nx = cn.x / nl; ny = cn.y / nl; nz = cn.z / nl; }
                float nc = rx * nx + ry * ny + rz * nz;
                float mn = fminf(nc, 0.0f);
                float px = rx - mn * nx, py = ry - mn * ny, pz = rz - mn * nz;
                float pl = len3(px, py, pz);
                if (nc < 0.0f && pl > 1e-20f) {
                    float sc = fmaxf(0.0f, pl + nc * col_friction) / pl;
                    px *= sc; py *= sc; pz *= sc;
                }
                vx = px + mx; vy = py + my; vz = pz + mz;
            }
        }
        // the mover accumulators are consumed (and cleared) even on steps without joint inputs
        if (mv_own.w != 0.0f || mv_own.x != 0.0f || mv_own.y != 0.0f || mv_own.z != 0.0f) g.mov[ni] = zero4;
        if (mv.w != 0.0f || mv.x != 0.0f || mv.y != 0.0f || mv.z != 0.0f) {
            if (use_mover && mv.w > 1e-15f) {
                float inv = 1.0f / mv.w;
                vx = mv.x * inv; vy = mv.y * inv; vz = mv.z * inv;
            }
        }
        if (n_bc > 0) {
            const int ix = ((co & 1023) << 2) + (l >> 4), iy = (((co >> 10) & 1023) << 2) + ((l >> 2) & 3),
                      iz = (((co >> 20) & 1023) << 2) + (l & 3);
            for (int k = 0; k < n_bc; k++) {
                const BCDesc& bc = bcs[k];
                const bool active = time >= bc.start_time && time < bc.end_time;
                if (bc.kind == 0) {
                    if (active) {
                        float ox = (float)ix * g.dx - bc.point[0], oy = (float)iy * g.dx - bc.point[1],
                              oz = (float)iz * g.dx - bc.point[2];
                        float dp = ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2];
                        if (dp < 0.0f) {
                            if (bc.surface_type == 11 && !((float)iz * g.dx < 0.4f || (float)iz * g.dx > 0.53f)) {
                                vx = vx * 0.3f; vy = 0.0f; vz = vz * 0.3f;
                            } else {
                                // sticky; the slip / friction branches also end in a zero store (:636-655)
                                vx = 0.f; vy = 0.f; vz = 0.f;
                            }
                        }
                    }
                } else if (bc.kind == 1) {
                    if (active) {
                        float ox = (float)ix * g.dx - bc.point[0], oy = (float)iy * g.dx - bc.point[1],
                              oz = (float)iz * g.dx - bc.point[2];
                        if (fabsf(ox) < bc.size[0] && fabsf(oy) < bc.size[1] && fabsf(oz) < bc.size[2]) {
                            vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
                        }
                    } else if (bc.reset == 1) {
                        if (time < bc.end_time + 15.0f * dt) { vx = 0.f; vy = 0.f; vz = 0.f; }
                    }
                } else if (bc.kind == 2) {
                    if (active) {
                        const int pad = 3;
                        if (ix < pad && vx < 0.f) vx = 0.f;
                        if (ix >= g.n - pad && vx > 0.f) vx = 0.f;
                        if (iy < pad && vy < 0.f) vy = 0.f;
                        if (iy >= g.n - pad && vy > 0.f) vy = 0.f;
                        if (iz < pad && vz < 0.f) vz = 0.f;
                        if (iz >= g.n - pad && vz > 0.f) vz = 0.f;
                    }
                } else if (bc.kind == 3) {
                    if (ix < g.n && iy < g.n && iz < g.n && bc.mask[((size_t)ix * g.n + iy) * g.n + iz] >= 1) {
                        vx = 0.f; vy = 0.f; vz = 0.f;
                    }
                }
            }
        }
        g.vout[ni] = make_float4(vx, vy, vz, 0.0f);
    }
    }  // pass
    if (PEER && g.ts) __syncthreads();  // probe only: the stamps are thread 0's, and other warps of the CTA may still be polling
    if (PEER) ts_end(g, TS_PULL);
    ts_end(g, TS_GRID);
    if (PEER) {  // the last CTA closes the exchange
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned done = atomicAdd(&P.counter[1], 1u);
            if (done == gridDim.x - 1) {
                P.counter[1] = 0;
                *P.epoch = E + 1;
            }
        }
    }
}
__global__ void k_reset_k(StepState* st) { st->k = 0; }
// per-call inputs of p2g2p (body points / velocities, joint velocities) from DEVICE pointers into the solver's own
// buffers, and the substep counter reset, in one launch
struct StageInputs {
    const float* src[5];
    float* dst[5];
    int n[5];  // floats; 0 = not given
};
__global__ void k_stage_inputs(StageInputs a, StepState* st) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->k = 0;
    const int stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 5; j++)
        for (int i = t; i < a.n[j]; i += stride) a.dst[j][i] = a.src[j][i];
}

// ============================================================ G2P
struct Gathered {
    float v[3];
    float C[9];
    float G[9];  // grad v
};
// Contraction of one 27-node tile (g2p_v / g2p_e, mpm_utils.py:726-763, 798-836) by sum factorisation:
// v = sum w v_n, C = 4/dx sum w v_n (x) (ijk - f), grad v = sum v_n (x) grad w are separable, so contract
// over k, then j, then i.  Per axis a:  W weight,  CW = W (i - f_a) 4/dx (APIC),  DW = dw / dx (gradient).
// Written on packed fp32 pairs (FFMA2): the x,y components of a node velocity ride in one pair with the
// weight broadcast, the z components of two different sums share a pair with a PAIR of weights.
// Outputs that the caller does not use are removed by the compiler.
__device__ __forceinline__ void g2p_contract(const float4* __restrict__ T, const float (&W)[3][3], const float (&CW)[3][3],
                                             const float (&DW)[3][3], Gathered& o) {
    const float2 z2 = make_float2(0.f, 0.f);
    float2 v_xy = z2, C0_xy = z2, C1_xy = z2, C2_xy = z2, G0_xy = z2, G1_xy = z2, G2_xy = z2;
    float2 vz_c1z = z2, c2z_g2z = z2, c0z_g0z = z2;
    float g1z = 0.f;
#ifdef MPM_G2P_ROWPF
    float4 nx[3];
#pragma unroll
    for (int q = 0; q < 3; q++) nx[q] = T[q];
#endif
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float2 A0_xy = z2, A1_xy = z2, A2_xy = z2, B0_xy = z2, C0q_xy = z2, A0z_A1z = z2, B0z_C0z = z2;
        float A2z = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            // one z-row (3 nodes, 12 registers) is fetched at a time; the compiler barrier keeps ptxas from
            // hoisting all 27 LDS.128 (108 registers) to the top and spilling
            float4 pl[3];
#ifdef MPM_G2P_ROWPF  // two z-rows in flight: the next row is fetched before the current one is consumed
#pragma unroll
            for (int q = 0; q < 3; q++) pl[q] = nx[q];
            if (i * 3 + j < 8) {
#pragma unroll
                for (int q = 0; q < 3; q++) nx[q] = T[(i * 3 + j + 1) * 3 + q];
            }
            asm volatile("" ::: "memory");
#else
#pragma unroll
            for (int q = 0; q < 3; q++) pl[q] = T[(i * 3 + j) * 3 + q];
            asm volatile("" ::: "memory");
#endif
            float2 a_xy = z2, b_xy = z2, c_xy = z2, bz_cz = z2;
            float az = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float4 gv = pl[k];
                const float2 gxy = make_float2(gv.x, gv.y);
                a_xy = fma2(W[2][k], gxy, a_xy);
                b_xy = fma2(CW[2][k], gxy, b_xy);
                c_xy = fma2(DW[2][k], gxy, c_xy);
                az = fmaf(W[2][k], gv.z, az);
                bz_cz = fma2(make_float2(CW[2][k], DW[2][k]), gv.z, bz_cz);
            }
            A0_xy = fma2(W[1][j], a_xy, A0_xy);
            A1_xy = fma2(CW[1][j], a_xy, A1_xy);
            A2_xy = fma2(DW[1][j], a_xy, A2_xy);
            B0_xy = fma2(W[1][j], b_xy, B0_xy);
            C0q_xy = fma2(W[1][j], c_xy, C0q_xy);
            A0z_A1z = fma2(make_float2(W[1][j], CW[1][j]), az, A0z_A1z);
            A2z = fmaf(DW[1][j], az, A2z);
            B0z_C0z = fma2(W[1][j], bz_cz, B0z_C0z);
        }
        v_xy = fma2(W[0][i], A0_xy, v_xy);
        C0_xy = fma2(CW[0][i], A0_xy, C0_xy);
        C1_xy = fma2(W[0][i], A1_xy, C1_xy);
        C2_xy = fma2(W[0][i], B0_xy, C2_xy);
        G0_xy = fma2(DW[0][i], A0_xy, G0_xy);
        G1_xy = fma2(W[0][i], A2_xy, G1_xy);
        G2_xy = fma2(W[0][i], C0q_xy, G2_xy);
        vz_c1z = fma2(W[0][i], A0z_A1z, vz_c1z);
        c2z_g2z = fma2(W[0][i], B0z_C0z, c2z_g2z);
        c0z_g0z = fma2(make_float2(CW[0][i], DW[0][i]), A0z_A1z.x, c0z_g0z);
        g1z = fmaf(W[0][i], A2z, g1z);
    }
    o.v[0] = v_xy.x; o.v[1] = v_xy.y; o.v[2] = vz_c1z.x;
    o.C[0] = C0_xy.x; o.C[3] = C0_xy.y; o.C[6] = c0z_g0z.x;
    o.C[1] = C1_xy.x; o.C[4] = C1_xy.y; o.C[7] = vz_c1z.y;
    o.C[2] = C2_xy.x; o.C[5] = C2_xy.y; o.C[8] = c2z_g2z.x;
    o.G[0] = G0_xy.x; o.G[3] = G0_xy.y; o.G[6] = c0z_g0z.y;
    o.G[1] = G1_xy.x; o.G[4] = G1_xy.y; o.G[7] = g1z;
    o.G[2] = G2_xy.x; o.G[5] = G2_xy.y; o.G[8] = c2z_g2z.y;
}

// Register-lean variants of the same contraction for the two cloth kernels.  The unrolled version above keeps the 27
// weights, three levels of partial sums and hoisted tile rows live (112-121 registers, 16 warps per SM); here the x
// axis is a ROLLED loop whose weights are evaluated in the loop, and only what the kernel uses is accumulated:
//   MODE 0 (vertices)  v and C;
//   MODE 1 (elements)  C and (grad v) d3 -- g2p_e only uses grad v in F d3 = d3 + dt (grad v) d3 (mpm_utils.py:850-855), so
//                      the three gradient weights are folded with d3 and ONE vector is accumulated instead of nine sums.
template <int MODE>
__device__ __forceinline__ void g2p_contract_lean(const float4* __restrict__ T, const float (&f)[3], float inv_dx, const float (&d3)[3],
                                                  Gathered& o) {
    float W1[3], CW1[3], W2[3], CW2[3], D1[3], D2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float ww, dd;
        bspline(f[1], i, ww, dd);
        W1[i] = ww; CW1[i] = ww * ((float)i - f[1]) * (inv_dx * 4.0f); D1[i] = dd * inv_dx * d3[1];
        bspline(f[2], i, ww, dd);
        W2[i] = ww; CW2[i] = ww * ((float)i - f[2]) * (inv_dx * 4.0f); D2[i] = dd * inv_dx * d3[2];
    }
    const float2 z2 = make_float2(0.f, 0.f);
    float2 v_xy = z2, C0_xy = z2, C1_xy = z2, C2_xy = z2, g_xy = z2;
    float vz = 0.f, c0z = 0.f, c1z = 0.f, c2z = 0.f, gz = 0.f;
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        // bspline(f[0], i) with a run-time i (same arithmetic as the unrolled form)
        const float fi = (float)i;
        const float A = (i == 1) ? -1.0f : 0.5f, B = (i == 1) ? 0.75f : 0.0f;
        const float t = f[0] - (1.5f - 0.5f * fi);
        const float w0 = A * t * t + B;
        const float cw0 = w0 * (fi - f[0]) * (inv_dx * 4.0f);
        const float d0 = (2.0f * A * t) * inv_dx * d3[0];
        const float4* __restrict__ Ti = T + i * 9;
        float2 A0_xy = z2, A1_xy = z2, B0_xy = z2, E_xy = z2;
        float A0z = 0.f, A1z = 0.f, B0z = 0.f, Ez = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float4 pl[3];
#pragma unroll
            for (int q = 0; q < 3; q++) pl[q] = Ti[j * 3 + q];
            asm volatile("" ::: "memory");  // one z-row in flight: keeps the loads from being hoisted into 100+ registers
            float2 a_xy = z2, b_xy = z2, c_xy = z2, bz_cz = z2;
            float az = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float2 gxy = make_float2(pl[k].x, pl[k].y);
                a_xy = fma2(W2[k], gxy, a_xy);
                b_xy = fma2(CW2[k], gxy, b_xy);
                az = fmaf(W2[k], pl[k].z, az);
                if (MODE == 1) {
                    c_xy = fma2(D2[k], gxy, c_xy);
                    bz_cz = fma2(make_float2(CW2[k], D2[k]), pl[k].z, bz_cz);
                } else {
                    bz_cz.x = fmaf(CW2[k], pl[k].z, bz_cz.x);
                }
            }
            A0_xy = fma2(W1[j], a_xy, A0_xy);
            A1_xy = fma2(CW1[j], a_xy, A1_xy);
            B0_xy = fma2(W1[j], b_xy, B0_xy);
            A0z = fmaf(W1[j], az, A0z);
            A1z = fmaf(CW1[j], az, A1z);
            B0z = fmaf(W1[j], bz_cz.x, B0z);
            if (MODE == 1) {  // E = sum_j (w1j c' + dw1j d3y a): the y and z terms of (grad w . d3) for this i
                E_xy = fma2(W1[j], c_xy, E_xy);
                E_xy = fma2(D1[j], a_xy, E_xy);
                Ez = fmaf(W1[j], bz_cz.y, Ez);
                Ez = fmaf(D1[j], az, Ez);
            }
        }
        if (MODE == 0) {
            v_xy = fma2(w0, A0_xy, v_xy);
            vz = fmaf(w0, A0z, vz);
        }
        C0_xy = fma2(cw0, A0_xy, C0_xy);
        C1_xy = fma2(w0, A1_xy, C1_xy);
        C2_xy = fma2(w0, B0_xy, C2_xy);
        c0z = fmaf(cw0, A0z, c0z);
        c1z = fmaf(w0, A1z, c1z);
        c2z = fmaf(w0, B0z, c2z);
        if (MODE == 1) {
            g_xy = fma2(w0, E_xy, g_xy);
            g_xy = fma2(d0, A0_xy, g_xy);
            gz = fmaf(w0, Ez, gz);
            gz = fmaf(d0, A0z, gz);
        }
    }
    o.v[0] = v_xy.x; o.v[1] = v_xy.y; o.v[2] = vz;
    o.C[0] = C0_xy.x; o.C[3] = C0_xy.y; o.C[6] = c0z;
    o.C[1] = C1_xy.x; o.C[4] = C1_xy.y; o.C[7] = c1z;
    o.C[2] = C2_xy.x; o.C[5] = C2_xy.y; o.C[8] = c2z;
    o.G[0] = g_xy.x; o.G[1] = g_xy.y; o.G[2] = gz;  // MODE 1: (grad v) d3
}

// Warp-collective gather.  The particles of a warp form a few same-cell runs; for each run lanes 0..26
// fetch the run's 27 node velocities ONCE (one table lookup + one LDG.128 per lane, up to G2P_RMAX runs
// in flight) into a shared-memory tile, then lane = particle contracts its run's tile with broadcast
// reads.  Replaces 8 table lookups + 27 dependent gathers per particle.
#ifndef MPM_G2P_RMAX
#define MPM_G2P_RMAX 16
#endif
constexpr int G2P_RMAX = MPM_G2P_RMAX;  // runs per contraction pass (tile capacity)
constexpr int G2P_OFF_B = 32 * 9 * 4;                        // nine per-axis node offsets for each of up to 32 runs
constexpr int G2P_TILE_B = G2P_OFF_B + G2P_RMAX * 27 * 16;  // run offsets | tiles
struct Gather {
    const Grid& g;
    const Warp& w;
    int* runoff;  // nine per-axis node offsets of every run (set by the run's leader lane)
    float4* tile;
    Runs R;
    bool valid;
    float f[3];
    int b[3];
    __device__ __forceinline__ Gather(const Grid& g_, const Warp& w_, unsigned char* tile_mem, bool valid_)
        : g(g_), w(w_), runoff(reinterpret_cast<int*>(tile_mem)), tile(reinterpret_cast<float4*>(tile_mem + G2P_OFF_B)), valid(valid_) {}
    // the packed stencil base of every lane's particle is all the staging needs: with the per-particle cell
    // arrays (CV / CE) the node loads are issued before the particle slab has arrived
    __device__ __forceinline__ void begin(int cell) {
        if (!valid) cell = -1 - w.lane;
        R = group_runs(w.lane, w.cnt, cell);  // one tile per DISTINCT cell, wherever its particles sit in the slab
        if ((R.starts >> w.lane) & 1u) stencil_offsets(g, cell, runoff + R.mine * 9);  // the group's leader decodes its cell once
        __syncwarp();
    }
    __device__ __forceinline__ int set_position(float x, float y, float z) {
        const float gp[3] = {x * g.inv_dx, y * g.inv_dx, z * g.inv_dx};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            b[a] = (int)(gp[a] - 0.5f);
            f[a] = gp[a] - (float)b[a];
        }
        return pack_cell(b[0], b[1], b[2]);
    }
    // lanes 0..26 copy the nodes of runs [r0, r0 + G2P_RMAX) into the tile with cp.async (LDGSTS): global ->
    // shared without staging registers, every run of the pass in flight at once.  A node's address is the sum of
    // the three per-axis offsets its run's leader has stored; a stencil that leaves the grid and the unused slots of the
    // last trip are zero-filled by a cp.async of source size 0 instead of a branch.
    __device__ __forceinline__ void stage_issue(int r0) const {
        const int nrp = min(G2P_RMAX, R.nr - r0);
        if (w.lane < 27) {
            const LaneNode ln(w.lane);
            float4* dst = tile + w.lane;
            for (int rb = 0; rb < nrp; rb += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    bool ok;
                    const int ni = ln.node(runoff + (r0 + min(rb + u, nrp - 1)) * 9, ok);
                    ok = ok && rb + u < nrp;
                    cp_async16_zfill(dst + (rb + u) * 27, g.vout + (ok ? ni : 0), ok ? 16 : 0);
                }
            }
        }
    }
    __device__ __forceinline__ void stage_wait() const {
        cp_async_wait_all();
        __syncwarp();
    }
    // the register-lean contractions (MODE 0: v, C; MODE 1: C, (grad v) d3)
    template <int MODE>
    __device__ __forceinline__ void contract_lean(int r0, const float (&d3)[3], Gathered& o) const {
        if (valid && R.mine >= r0 && R.mine < r0 + G2P_RMAX) g2p_contract_lean<MODE>(tile + (R.mine - r0) * 27, f, g.inv_dx, d3, o);
    }
    template <int MODE>
    __device__ __forceinline__ void remaining_passes_lean(const float (&d3)[3], Gathered& o) const {
        for (int r0 = G2P_RMAX; r0 < R.nr; r0 += G2P_RMAX) {
            __syncwarp();
            stage_issue(r0);
            stage_wait();
            contract_lean<MODE>(r0, d3, o);
        }
    }
    // lane = particle: contract the tile of my run if it is staged
    __device__ __forceinline__ void contract(int r0, Gathered& o) const {
        if (valid && R.mine >= r0 && R.mine < r0 + G2P_RMAX) {
            float W[3][3], CW[3][3], DW[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    float ww, dd;
                    bspline(f[a], i, ww, dd);
                    W[a][i] = ww;
                    CW[a][i] = ww * ((float)i - f[a]) * (g.inv_dx * 4.0f);
                    DW[a][i] = dd * g.inv_dx;
                }
            g2p_contract(tile + (R.mine - r0) * 27, W, CW, DW, o);
        }
    }
    // passes beyond the first (a warp spanning more than G2P_RMAX cells): stage, then contract
    __device__ __forceinline__ void remaining_passes(Gathered& o) const {
        for (int r0 = G2P_RMAX; r0 < R.nr; r0 += G2P_RMAX) {
            __syncwarp();
            stage_issue(r0);
            stage_wait();
            contract(r0, o);
        }
    }
};
__device__ __forceinline__ float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
// the blocks under the particle's new stencil exist already unless its base cell changed
__device__ __forceinline__ void ensure_if_moved(const Grid& g, const int (&b)[3], float x, float y, float z) {
    if (base_of(x, g.inv_dx) != b[0] || base_of(y, g.inv_dx) != b[1] || base_of(z, g.inv_dx) != b[2]) ensure_stencil_blocks(g, x, y, z);
}

// end of substep: self.time += dt (mpm_solver.py:536), substep counter, moving cuboids (:975-981);
// run by one thread of the LAST kernel of the substep
__device__ __forceinline__ void advance_step(StepState* st, float dt, BCDesc* bcs, int n_bc) {
    float time = (float)st->time;
    for (int k = 0; k < n_bc; k++)
        if (bcs[k].kind == 1 && time >= bcs[k].start_time && time < bcs[k].end_time)
            for (int a = 0; a < 3; a++) bcs[k].point[a] = bcs[k].point[a] + dt * bcs[k].velocity[a];
    st->time = st->time + (double)dt;
    st->k = st->k + 1;
}
struct Advance {
    StepState* st;  // null: this kernel is not the last one of the substep
    BCDesc* bcs;
    int n_bc;
};

#ifndef MPM_G2P_NW
#define MPM_G2P_NW 1
#endif
constexpr int G2P_NW = MPM_G2P_NW;  // warps (= slabs) per CTA
constexpr int G2P_MINB = 16 / G2P_NW;  // unrolled contraction (traditional particles): 16 resident warps per SM, <= 128 registers
#ifndef MPM_G2P_V_WARPS
#define MPM_G2P_V_WARPS 20
#endif
#ifndef MPM_G2P_E_WARPS
#define MPM_G2P_E_WARPS 20
#endif
constexpr int G2P_V_MINB = MPM_G2P_V_WARPS / G2P_NW, G2P_E_MINB = MPM_G2P_E_WARPS / G2P_NW;  // lean contractions
// The gather-side kernels run in the order  grid update -> vertex G2P -> [traditional G2P] -> element G2P.
// Only the FIRST of them waits for the grid update (wait_first); a follower starts when every CTA of its predecessor
// has passed that wait (PDL invariant 2), so the grid velocities are final for it too: it works while its predecessor
// drains, lets its successor in when its own gather is done, and waits before it exits (invariant 1) -- the element
// kernel earlier: in front of its corner reads, which need the moved vertices.
// g2p_v for cloth vertices (mpm_utils.py:716-786); also clears the vertex_force accumulator of the next substep
// and allocates grid blocks for the new position.
constexpr int G2P_V_WB = VP_F * 32 * 4 + G2P_TILE_B;
__global__ void __launch_bounds__(32 * G2P_NW, G2P_V_MINB) k_g2p_vertices(Grid g, int Nv, float* __restrict__ VP, float4* __restrict__ VFnext,
                                                               int* __restrict__ CV, float dt, int wait_first, Advance adv) {
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<G2P_NW, G2P_V_WB>(w, Nv, smem)) return;
    ts_begin(g, TS_G2P_V);
    PHASE_BEGIN();
    float* sP = reinterpret_cast<float*>(w.buf);
    // VP and CV were last written by the previous substep's G2P: the slab load and the run search overlap the
    // tail of the grid update; only the node velocities need it
    slab_issue(w, {Slab{sP, VP, VP_F}});
    const bool valid = w.lane < w.cnt;
    const int p = w.p0 + w.lane;
    Gather G(g, w, w.buf + VP_F * 32 * 4, valid);
    G.begin(valid ? CV[p] : 0);
    if (wait_first) {
        pdl_wait();
        pdl_trigger();
    }
    G.stage_issue(0);  // node loads are in flight before the slab lands
    PHASE(g, 3, 1);
    slab_wait(w);
    PHASE(g, 3, 0);
    float4* r4 = reinterpret_cast<float4*>(sP + w.lane * VP_F);
    float4 xm = valid ? r4[0] : make_float4(0.f, 0.f, 0.f, 0.f);
    G.set_position(xm.x, xm.y, xm.z);
    G.stage_wait();
    PHASE(g, 3, 2);
    Gathered o;
    {
        const float no_d3[3] = {0.f, 0.f, 0.f};
        G.contract_lean<0>(0, no_d3, o);
        G.remaining_passes_lean<0>(no_d3, o);
    }
    PHASE(g, 3, 3);
    if (!wait_first) pdl_trigger();
    if (valid) {
        const float dxc = 1.0f / g.inv_dx, a_min = dxc * 2.0f, a_max = g.lim - dxc * 2.0f;
        xm.x = clampf(xm.x + dt * o.v[0], a_min, a_max);
        xm.y = clampf(xm.y + dt * o.v[1], a_min, a_max);
        xm.z = clampf(xm.z + dt * o.v[2], a_min, a_max);
        r4[0] = xm;
        r4[1] = make_float4(o.v[0], o.v[1], o.v[2], o.C[0]);
        r4[2] = make_float4(o.C[1], o.C[2], o.C[3], o.C[4]);
        r4[3] = make_float4(o.C[5], o.C[6], o.C[7], o.C[8]);
        // vertex forces are double buffered: the NEXT substep's accumulator is cleared here (replaces set_vec3_to_zero,
        // mpm_solver.py:251-256), this substep's forces stay readable as state.vertex_force
        VFnext[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int nb0 = base_of(xm.x, g.inv_dx), nb1 = base_of(xm.y, g.inv_dx), nb2 = base_of(xm.z, g.inv_dx);
        if (nb0 != G.b[0] || nb1 != G.b[1] || nb2 != G.b[2]) {  // the blocks under the new stencil exist unless the cell changed
            ensure_stencil_blocks(g, xm.x, xm.y, xm.z);
            CV[p] = pack_cell(nb0, nb1, nb2);
        }
    }
    PHASE(g, 3, 4);
    slab_store(w, {Slab{sP, VP, VP_F}});
    PHASE(g, 3, 5);
    PHASE_END(g, 3);
    ts_end(g, TS_G2P_V);
    if (!wait_first) pdl_wait();
    if (adv.st && blockIdx.x == 0 && threadIdx.x == 0) advance_step(adv.st, dt, adv.bcs, adv.n_bc);  // every predecessor has completed
}

// g2p_v for traditional particles: additionally F_trial = (I + dt grad v) F (mpm_utils.py:783-786)
constexpr int G2P_T_WB = (KP_F + TF_F) * 32 * 4 + G2P_TILE_B;
__global__ void __launch_bounds__(32 * G2P_NW, G2P_MINB) k_g2p_traditional(Grid g, int Nt, float* __restrict__ TP, float* __restrict__ TF,
                                                                  float dt, int wait_first, Advance adv) {
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<G2P_NW, G2P_T_WB>(w, Nt, smem)) return;
    ts_begin(g, TS_G2P_T);
    float* sP = reinterpret_cast<float*>(w.buf);
    float* sT = sP + 32 * KP_F;
    if (wait_first) {
        pdl_wait();
        pdl_trigger();
    }
    slab_load(w, {Slab{sP, TP, KP_F}, Slab{sT, TF, TF_F}});
    const bool valid = w.lane < w.cnt;
    float* r = sP + w.lane * KP_F;
    float x = valid ? r[0] : 0.f, y = valid ? r[1] : 0.f, z = valid ? r[2] : 0.f;
    Gathered o;
    Gather G(g, w, w.buf + (KP_F + TF_F) * 32 * 4, valid);
    G.begin(G.set_position(x, y, z));
    G.stage_issue(0);
    G.stage_wait();
    G.contract(0, o);
    G.remaining_passes(o);
    if (!wait_first) pdl_trigger();
    if (valid) {
        float* t = sT + w.lane * TF_F;
        const float dxc = 1.0f / g.inv_dx, a_min = dxc * 2.0f, a_max = g.lim - dxc * 2.0f;
        x = clampf(x + dt * o.v[0], a_min, a_max);
        y = clampf(y + dt * o.v[1], a_min, a_max);
        z = clampf(z + dt * o.v[2], a_min, a_max);
        r[0] = x; r[1] = y; r[2] = z;
        r[P_V] = o.v[0]; r[P_V + 1] = o.v[1]; r[P_V + 2] = o.v[2];
#pragma unroll
        for (int i = 0; i < 9; i++) r[P_C + i] = o.C[i];
        float M[9], F[9], Ft[9];
#pragma unroll
        for (int i = 0; i < 9; i++) { M[i] = o.G[i] * dt; F[i] = t[T_F + i]; }
        M[0] += 1.0f; M[4] += 1.0f; M[8] += 1.0f;
        mat_mul(M, F, Ft);
#pragma unroll
        for (int i = 0; i < 9; i++) t[T_FT + i] = Ft[i];
        ensure_if_moved(g, G.b, x, y, z);
    }
    slab_store(w, {Slab{sP, TP, KP_F}, Slab{sT, TF, TF_F}});
    ts_end(g, TS_G2P_T);
    if (!wait_first) pdl_wait();
    if (adv.st && blockIdx.x == 0 && threadIdx.x == 0) advance_step(adv.st, dt, adv.bcs, adv.n_bc);
}

// g2p_e (mpm_utils.py:788-857): C and grad v at the OLD centroid, x/v = mean of the three
// already-updated corner vertices, d = [x2-x1, x3-x1, (I + dt grad v) d3].  All element state moves as coalesced
// float4 streams; the only shared memory is the node tile of the gather.  Reads the return-mapped d3 of direction
// buffer `cur`, writes buffer `cur^1`.  A follower of the vertex G2P: gather + contraction need the grid update only,
// griddepcontrol.wait sits in front of the corner reads.
struct ElemG2P {
    const int4* EFM;
    float4 *XE, *EV, *ED1, *ED2, *C0, *C1;
    const float4* D3in;
    float4* D3out;
    int* CE;
    const float* VP;
};
constexpr int G2P_E_WB = G2P_TILE_B + 512;  // + the corner slots of the 32 elements, parked by cp.async until the contraction is done
__global__ void __launch_bounds__(32 * G2P_NW, G2P_E_MINB) k_g2p_elements(Grid g, int Ne, ElemG2P A, float dt, Advance adv) {
    extern __shared__ __align__(128) unsigned char smem[];
    Warp w;
    if (!warp_begin<G2P_NW, G2P_E_WB>(w, Ne, smem)) return;
    ts_begin(g, TS_G2P_E);
    PHASE_BEGIN();
    const bool valid = w.lane < w.cnt;
    const int p = w.p0 + w.lane;
    const int cell = valid ? A.CE[p] : 0;
    const float4 xe = A.XE[p], d3v4 = A.D3in[p];
    Gather G(g, w, w.buf, valid);
    G.begin(cell);
    // the corner slots are needed after the contraction only: global -> shared without a register (LDGSTS), in flight
    // together with the node tiles, so that no load sits between the contraction and the corner reads
    int4* park = reinterpret_cast<int4*>(w.buf + G2P_TILE_B) + w.lane;
    cp_async16(park, &A.EFM[p]);
    G.stage_issue(0);
    PHASE(g, 5, 1);
    G.set_position(valid ? xe.x : 0.f, valid ? xe.y : 0.f, valid ? xe.z : 0.f);
    G.stage_wait();
    PHASE(g, 5, 2);
    {
        Gathered o;
        const float d3v[3] = {d3v4.x, d3v4.y, d3v4.z};
        G.contract_lean<1>(0, d3v, o);
        G.remaining_passes_lean<1>(d3v, o);
        if (valid) {
            A.C0[p] = make_float4(o.C[0], o.C[1], o.C[2], o.C[3]);
            A.C1[p] = make_float4(o.C[4], o.C[5], o.C[6], o.C[7]);
            // d3 <- (I + dt grad v) d3 (mpm_utils.py:850-855), with (grad v) d3 accumulated directly; C[8] rides along
            A.D3out[p] = make_float4(fmaf(dt, o.G[0], d3v[0]), fmaf(dt, o.G[1], d3v[1]), fmaf(dt, o.G[2], d3v[2]), o.C[8]);
        }
    }
    PHASE(g, 5, 4);  // contraction
    const int4 efm = *park;
    pdl_wait();      // the corner vertices must have been moved by the vertex G2P
    pdl_trigger();
    if (valid) {
        const float4* q0 = reinterpret_cast<const float4*>(A.VP + (size_t)efm.x * VP_F);
        const float4* q1 = reinterpret_cast<const float4*>(A.VP + (size_t)efm.y * VP_F);
        const float4* q2 = reinterpret_cast<const float4*>(A.VP + (size_t)efm.z * VP_F);
        const float4 x1 = q0[0], v1 = q0[1], x2 = q1[0], v2 = q1[1], x3 = q2[0], v3 = q2[1];
        const float nx = (x1.x + x2.x + x3.x) / 3.0f, ny = (x1.y + x2.y + x3.y) / 3.0f, nz = (x1.z + x2.z + x3.z) / 3.0f;
        A.XE[p] = make_float4(nx, ny, nz, 0.f);
        A.EV[p] = make_float4((v1.x + v2.x + v3.x) / 3.0f, (v1.y + v2.y + v3.y) / 3.0f, (v1.z + v2.z + v3.z) / 3.0f, 0.f);
        A.ED1[p] = make_float4(x2.x - x1.x, x2.y - x1.y, x2.z - x1.z, 0.f);
        A.ED2[p] = make_float4(x3.x - x1.x, x3.y - x1.y, x3.z - x1.z, 0.f);
        const int nb0 = base_of(nx, g.inv_dx), nb1 = base_of(ny, g.inv_dx), nb2 = base_of(nz, g.inv_dx);
        if (nb0 != G.b[0] || nb1 != G.b[1] || nb2 != G.b[2]) {
            ensure_stencil_blocks(g, nx, ny, nz);
            A.CE[p] = pack_cell(nb0, nb1, nb2);
        }
    }
    PHASE(g, 5, 3);  // corner reads
    PHASE_END(g, 5);
    ts_end(g, TS_G2P_E);
    if (adv.st && blockIdx.x == 0 && threadIdx.x == 0) advance_step(adv.st, dt, adv.bcs, adv.n_bc);  // every predecessor has completed
}

}  // namespace mpm
