// Device-side data layout and helpers of the B200 MPM substep solver.
//
// GRID: 4x4x4-node blocks, DIRECTLY ADDRESSED, SPARSELY VISITED.  Node (ix,iy,iz) lives at
// ((bx*nb+by)*nb+bz)*64 + local in every per-node array, which is SEPARABLE: index = X(ix) + Y(iy) + Z(iz) with one
// small term per axis (axis_offset), so the three offsets per axis of a cell run's stencil are computed once by the
// run's leader lane and every stencil node is a three-term sum (256^3: 268 MB per array, 512^3: 2.1 GB; 180 GB of
// HBM makes dense ADDRESSING affordable).  A block's 64 nodes are 1 KB of contiguous memory -- neighbouring cells
// share cache lines in all three directions (a plain x-major layout was measured: cheaper index arithmetic, but the
// gather and the grid update lose more to the scattered 48-byte rows than the arithmetic saves).  Only the
// ACTIVE blocks (those under some particle stencil) are ever touched: a block table (nb^3
// ints, -1 = inactive, else position in the active list slot_coord[]) is maintained by the kernels that
// move particles, and the grid update walks that list.  Nothing ever sweeps the dense arrays.
// Every per-node quantity is a float4 so that one P2G contribution is ONE 16-byte REDG.E.ADD.F32x4
// (sm_90+ vector atomic):
//   acc  {m*vx, m*vy, m*vz, m}      <- P2G                      (mpm_utils.py:548-557)
//   vout {vx, vy, vz, -}            <- grid update, read by G2P (mpm_utils.py:561-572)
//   colv {w*vx, w*vy, w*vz, w}, coln {w*nx, w*ny, w*nz, -}  <- body-mesh collider scatter
//                                                              (mpm_solver.py:829-880)
//   mov  {w*vx, w*vy, w*vz, w}      <- particle-mover scatter  (mpm_solver.py:677-788)
// Invariants: (1) every block under the stencil of any particle is in the active list; (2) between
// substeps all accumulators are zero (the grid update re-zeroes what it consumed), so there is no
// zero_grid sweep.
//
// PARTICLES: three classes (elements, traditional, vertices), each sorted by
// (Morton(block), cell-in-block).  A warp owns 32 consecutive particles of a class.
// Cloth ELEMENTS are a handful of float4 / int4 streams (lane = particle: every load / store is one fully coalesced
// LDG.128 / STG.128, no shared-memory staging):
//   EFM    int4   {corner vertex slots (sorted order) x3, mass as float bits}     read-only
//   K0,K1  float4 {Rinv[3], mu}, {lam, gamma, kappa, vol}                         read-only
//   XE,EV  float4 centroid / velocity = mean of the corner vertices               written by the element G2P
//   ED1,ED2 float4 in-plane directions d1, d2 = two edges                          written by the element G2P
//   CE     int    packed stencil base cell of XE               (lets G2P start its node loads early)
//   C0,C1  float4 {C[0..3]}, {C[4..7]}; C[8] rides in D3.w     written by the element G2P
//   D3     float4 {d3[3], C[8]}   ping-pong: return-mapped in place by the stress + P2G kernel (buffer `cur`),
//                                 advanced by G2P into buffer `cur^1`
//   SP3    float4 {vol * P3}      third stress column of the last substep (state.particle_stress = SP3 (x) d3)
// Traditional particles and vertices keep small AoS sub-records GROUPED BY THE KERNEL THAT WRITES THEM, moved
// with a single cp.async.bulk (TMA, SASS UBLKCP) per slab into / out of shared memory:
//   TP     kinematics  {x,y,z,m, vx,vy,vz,vol, C[9]}      17 floats  written by G2P
//   TS     stress      {S[9]}                              9 floats  written by the traditional stress kernel
//   TF     trad state  {F[9], Ft[9], mu, lam, ys}         21 floats  written by stress (F,..) and G2P (Ft)
//   VP     kinematics  {x,y,z,m, vx,vy,vz, C[9], pad[4]}  20 floats  written by G2P
//   VF     force       float4 {fx,fy,fz,-}   REDG.128 by stress; ping-pong like D3: G2P clears the buffer of the
//                                            NEXT substep, the last one stays readable (state.vertex_force)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpm {

constexpr int BS = 4;    // block edge in nodes
constexpr int BN = 64;   // nodes per block
constexpr int MAX_BC = 16;
constexpr int MAX_OPS = 64;

// sub-record sizes in floats
constexpr int KP_F = 17;  // TP
constexpr int S_F = 9;    // TS
constexpr int TF_F = 21;
constexpr int VP_F = 20;  // 16 used + 4 pad: an 80-byte stride makes lane = particle LDS.128 / STS.128 bank-conflict free
                          // (at 64 bytes they were 4-way conflicted: 64 of the ~290 shared-memory wavefronts of a G2P warp)
constexpr int VF_F = 4;
// field offsets
constexpr int P_X = 0, P_M = 3, P_V = 4, P_VOL = 7, P_C = 8;  // TP
constexpr int V_X = 0, V_M = 3, V_V = 4, V_C = 7;              // VP
constexpr int T_F = 0, T_FT = 9, T_MU = 18, T_LAM = 19, T_YS = 20;

struct Grid {
    int n, nb;
    float dx, inv_dx, lim;
    int cap;          // nb^3 blocks
    int* table;       // [nb^3] -1 = inactive, else index into slot_coord
    int* n_slots;     // device counter: number of active blocks
    int* slot_coord;  // [cap] active list: bx | by<<10 | bz<<20
    float4 *acc, *vout, *colv, *coln, *mov;
    float4* dbg_acc;  // copy of acc taken by the grid update when debugging (else null)
    int* flags;       // [0] pool overflow, [1] scatter/gather hit an unallocated block
    unsigned long long* clk;  // phase-clock accumulators (only read by -DMPM_CLK builds, tools/phase_clocks.py)
    unsigned long long* ts;   // timeline probe (mpm_measure_timeline): [kernel id][first start, last end] in ns, else null
};
// Timeline probe: first CTA start / last CTA end of a kernel on the GPU's global timer.  CUDA events cannot
// time kernels that overlap under programmatic dependent launch; these stamps can.  One 64-bit atomic per warp
// at each end, only when g.ts is set.
enum { TS_P2G_E = 0, TS_P2G_T, TS_P2G_V, TS_SCATTER, TS_GRID, TS_G2P_V, TS_G2P_T, TS_G2P_E, TS_KERNELS,
       TS_PUSH = TS_KERNELS, TS_PULL, TS_KERNELS_SHARDED };  // the two exchange kernels of a sharded substep
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// CTAs start in blockIdx order, so the first start is among the first CTAs and the last end (almost surely) among
// the last ones: only the first / last 512 CTAs stamp, which keeps the probe's same-address atomics off the
// measured path (stamping every warp cost ~10 % of a substep).
__device__ __forceinline__ void ts_begin(const Grid& g, int k) {
    if (g.ts && threadIdx.x == 0 && blockIdx.x < 512) atomicMin(&g.ts[2 * k], globaltimer_ns());
}
__device__ __forceinline__ void ts_end(const Grid& g, int k) {
    if (g.ts && threadIdx.x == 0 && blockIdx.x + 512 >= gridDim.x) atomicMax(&g.ts[2 * k + 1], globaltimer_ns());
}

// Per-warp phase clocks for latency analysis: PHASE_BEGIN at kernel entry, PHASE(k, i) after phase i of
// kernel k adds the elapsed SM cycles to a per-thread accumulator, PHASE_END(k) flushes lane 0's sums to
// clk[copy*64 + k*8 + i] (and counts warps in slot 7).
#ifdef MPM_CLK
#define PHASE_BEGIN() long long clk_prev_ = clock64(); unsigned clk_acc_[7] = {0, 0, 0, 0, 0, 0, 0}
#define PHASE(gr, k, i)                               \
    do {                                              \
        long long t_ = clock64();                     \
        clk_acc_[i] += (unsigned)(t_ - clk_prev_);    \
        clk_prev_ = t_;                               \
    } while (0)
// one flush per warp, spread over 64 copies of the table to keep the probe's own atomics cheap
#define PHASE_END(gr, k)                                                                                   \
    do {                                                                                                   \
        if ((threadIdx.x & 31) == 0) {                                                                     \
            unsigned long long* c_ = (gr).clk + (blockIdx.x & 63) * 64 + (k) * 8;                          \
            for (int i_ = 0; i_ < 7; i_++) if (clk_acc_[i_]) atomicAdd(&c_[i_], (unsigned long long)clk_acc_[i_]); \
            atomicAdd(&c_[7], 1ull);                                                                       \
        }                                                                                                  \
    } while (0)
#else
#define PHASE_BEGIN() do {} while (0)
#define PHASE(gr, k, i) do {} while (0)
#define PHASE_END(gr, k) do {} while (0)
#endif

// Peer-to-peer exchange of the grid blocks shared with other ranks (sharded runs; fused into k_grid_update<true>).
// Every rank owns a receive area, mapped into every peer (CUDA IPC):
//   [2 epoch parities][nranks senders][capA + capM blocks][64 nodes] x 32 bytes {x, flag, y, flag | z, flag, w, flag}
// A part is valid in epoch E when all four flags read E + 1 (the area starts zeroed; epochs only grow).  Every 8-byte
// {float, flag} pair is written by one store and read by one load, which the memory system never splits -- the "LL"
// protocol of NCCL: data and its arrival flag travel together, so there is no fence and no separate signal.
struct PeerArea {
    unsigned char* base[8];  // receive area of every rank as mapped here (base[rank] is the local one)
    unsigned long long slot_bytes;  // one sender's part of one parity: (capA + capM) * 64 * 32
    int rank, nranks;  // nranks == 0: not a peer-to-peer sharded step
    unsigned* epoch;    // completed exchanges (device)
    unsigned* counter;  // last-CTA detection of the grid update
    const int *listA, *listM;          // packed coordinates of the shared blocks: list A (acc) / M (mov)
    const int *nA, *nM;                // their lengths (device)
    const unsigned char *memA, *memM;  // member ranks of every listed block (bit r = rank r can reach it)
    const int *mapA, *mapM;            // block (table index) -> position in list A / M, -1 if not listed
    int capA, capM;                    // list capacities: list M's slots follow list A's in the receive areas
};
__device__ __forceinline__ void ll_store(unsigned char* p, const float4& v, unsigned flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(__float_as_uint(v.x)), "r"(flag), "r"(__float_as_uint(v.y)) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p + 16), "r"(__float_as_uint(v.z)), "r"(flag), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ float4 ll_load(const unsigned char* p, unsigned flag) {
    unsigned x, fx, y, fy, z, fz, w, fw;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(fx), "=r"(y), "=r"(fy) : "l"(p) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(z), "=r"(fz), "=r"(w), "=r"(fw) : "l"(p + 16) : "memory");
    } while (fx != flag || fy != flag || fz != flag || fw != flag);
    return make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(w));
}

struct StepState {
    double time;  // MPMWARP.time (mpm_solver.py:28,536)
    int k;        // substep index inside the current mpm_step call
    int pad;
};

struct BCDesc {  // one grid_postprocess entry (mpm_solver.py:564-658, 929-1053, 1330-1355)
    int kind;    // 0 surface, 1 cuboid, 2 bounding box, 3 mask
    int surface_type, reset, pad;
    float point[3], normal[3], size[3], velocity[3];
    float friction, start_time, end_time;
    const int* mask;
};

struct ParticleOp {  // pre-P2G operations (mpm_solver.py:1058-1328, 1360-1417)
    int kind;        // 0 impulse/mass, 1 impulse, 2 set velocity, 3 rotate about an axis (cylinder selection)
    float vec[3];
    float start_time, end_time;
    const int* mask;  // canonical order [N]
    // kind 3 (enforce_particle_velocity_rotation, :1156-1256): v = -h sin(t) rot * h1 + h cos(t) rot * h2 + trans * n
    float point[3], n[3], h1[3], h2[3], rot, trans;
};

struct ModelDev {
    int material, hardening;
    float friction_coeff, alpha;
    float gx, gy, gz;
    float rpic, damping, xi, plastic_viscosity, softening;
};

// ------------------------------------------------------------------ small math
__device__ __forceinline__ void mat_mul(const float* a, const float* b, float* o) {
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) o[i] = t[i];
}
// o = a * b^T
__device__ __forceinline__ void mat_mul_bt(const float* a, const float* b, float* o) {
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
#pragma unroll
    for (int i = 0; i < 9; i++) o[i] = t[i];
}
__device__ __forceinline__ float det3(const float* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ __forceinline__ float len3(float a, float b, float c) { return sqrtf(a * a + b * b + c * c); }
// un-contracted (no FMA) dot / length: the cloth return mapping branches on R22 > 1 exactly at the
// rest state (mpm_utils.py:196), so the QR that feeds it is evaluated in plain IEEE fp32
// operation order to keep the branch reproducible against a non-FMA evaluation
__device__ __forceinline__ float dot3_rn(const float* a, const float* b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}
__device__ __forceinline__ float len3_rn(const float* a) { return __fsqrt_rn(dot3_rn(a, a)); }

// Jacobi rotation on the symmetric matrix (bpp,bqq,bpq; brp,brq with r the third index) and V columns p,q
__device__ __forceinline__ void jrot(float& bpp, float& bqq, float& bpq, float& brp, float& brq, float* V, int p, int q) {
    if (fabsf(bpq) < 1e-30f) return;
    float theta = (bqq - bpp) / (2.0f * bpq);
    float t = copysignf(1.0f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
    float c = rsqrtf(t * t + 1.0f), s = t * c;
    bpp -= t * bpq;
    bqq += t * bpq;
    bpq = 0.0f;
    float rp = brp, rq = brq;
    brp = c * rp - s * rq;
    brq = s * rp + c * rq;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float vp = V[3 * r + p], vq = V[3 * r + q];
        V[3 * r + p] = c * vp - s * vq;
        V[3 * r + q] = s * vp + c * vq;
    }
}
__device__ __forceinline__ void swap_cols(float* V, int a, int b, float& la, float& lb) {
    // swap two columns and negate one to keep det V = +1
    float t = la; la = lb; lb = t;
#pragma unroll
    for (int r = 0; r < 3; r++) { float x = V[3 * r + a]; V[3 * r + a] = V[3 * r + b]; V[3 * r + b] = -x; }
}
// A = U diag(S) V^T, det U = det V = +1, S[0]>=S[1]>=|S[2]| (wp.svd3 convention; stands in for
// warp-lang's McAdams solver at mpm_utils.py:217,265,322,369,1077)
__device__ __forceinline__ void svd3(const float* A, float* U, float* S, float* V) {
    float b00 = A[0] * A[0] + A[3] * A[3] + A[6] * A[6];
    float b11 = A[1] * A[1] + A[4] * A[4] + A[7] * A[7];
    float b22 = A[2] * A[2] + A[5] * A[5] + A[8] * A[8];
    float b01 = A[0] * A[1] + A[3] * A[4] + A[6] * A[7];
    float b02 = A[0] * A[2] + A[3] * A[5] + A[6] * A[8];
    float b12 = A[1] * A[2] + A[4] * A[5] + A[7] * A[8];
    V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1;
#pragma unroll 1
    for (int sweep = 0; sweep < 6; sweep++) {
        jrot(b00, b11, b01, b02, b12, V, 0, 1);
        jrot(b00, b22, b02, b01, b12, V, 0, 2);
        jrot(b11, b22, b12, b01, b02, V, 1, 2);
    }
    if (b00 < b11) swap_cols(V, 0, 1, b00, b11);
    if (b00 < b22) swap_cols(V, 0, 2, b00, b22);
    if (b11 < b22) swap_cols(V, 1, 2, b11, b22);
    float av[9];  // av[3*r+c] = (A V)[r][c]
    mat_mul(A, V, av);
    float s0 = len3(av[0], av[3], av[6]);
    float u0[3], u1[3], u2[3];
    if (s0 > 1e-30f) { float i0 = 1.0f / s0; u0[0] = av[0] * i0; u0[1] = av[3] * i0; u0[2] = av[6] * i0; }
    else { u0[0] = 1; u0[1] = 0; u0[2] = 0; }
    float dp = u0[0] * av[1] + u0[1] * av[4] + u0[2] * av[7];
    float w0 = av[1] - dp * u0[0], w1 = av[4] - dp * u0[1], w2 = av[7] - dp * u0[2];
    float n1 = len3(w0, w1, w2);
    if (n1 > 1e-6f * fmaxf(s0, 1e-30f)) { float i1 = 1.0f / n1; u1[0] = w0 * i1; u1[1] = w1 * i1; u1[2] = w2 * i1; }
    else {
        float a0 = fabsf(u0[0]), a1 = fabsf(u0[1]), a2 = fabsf(u0[2]);
        float e0 = (a0 <= a1 && a0 <= a2) ? 1.f : 0.f, e1 = (e0 == 0.f && a1 <= a2) ? 1.f : 0.f, e2 = 1.f - e0 - e1;
        float dd = e0 * u0[0] + e1 * u0[1] + e2 * u0[2];
        w0 = e0 - dd * u0[0]; w1 = e1 - dd * u0[1]; w2 = e2 - dd * u0[2];
        float i1 = rsqrtf(w0 * w0 + w1 * w1 + w2 * w2);
        u1[0] = w0 * i1; u1[1] = w1 * i1; u1[2] = w2 * i1;
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    S[0] = s0;
    S[1] = u1[0] * av[1] + u1[1] * av[4] + u1[2] * av[7];
    S[2] = u2[0] * av[2] + u2[1] * av[5] + u2[2] * av[8];
#pragma unroll
    for (int r = 0; r < 3; r++) { U[3 * r] = u0[r]; U[3 * r + 1] = u1[r]; U[3 * r + 2] = u2[r]; }
}
// o = U diag(dg) V^T
__device__ __forceinline__ void diag_sandwich(const float* U, const float* dg, const float* V, float* o) {
    float ud[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) ud[3 * r + c] = U[3 * r + c] * dg[c];
    mat_mul_bt(ud, V, o);
}

// ------------------------------------------------------------------ grid helpers
__device__ __forceinline__ uint32_t part1by2(uint32_t x) {
    x &= 0x3ff;
    x = (x | (x << 16)) & 0x30000ff;
    x = (x | (x << 8)) & 0x300f00f;
    x = (x | (x << 4)) & 0x30c30c3;
    x = (x | (x << 2)) & 0x9249249;
    return x;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// wp.int(x*inv_dx - 0.5) (truncation toward zero), mpm_utils.py:499-502
__device__ __forceinline__ int base_of(float x, float inv_dx) { return (int)(x * inv_dx - 0.5f); }

__device__ __forceinline__ uint32_t sort_key(const Grid& g, float x, float y, float z) {
    int bx = clampi(base_of(x, g.inv_dx), 0, g.n - 1), by = clampi(base_of(y, g.inv_dx), 0, g.n - 1),
        bz = clampi(base_of(z, g.inv_dx), 0, g.n - 1);
    uint32_t m = part1by2(bx >> 2) | (part1by2(by >> 2) << 1) | (part1by2(bz >> 2) << 2);
    return (m << 6) | ((bx & 3) << 4) | ((by & 3) << 2) | (bz & 3);
}
__device__ __forceinline__ int table_index(const Grid& g, int bx, int by, int bz) { return (bx * g.nb + by) * g.nb + bz; }

// >= 0 iff the block is in the active list (plain load: G2P activates blocks in the same kernel)
__device__ __forceinline__ int lookup_slot(const Grid& g, int bx, int by, int bz) {
    return g.table[table_index(g, bx, by, bz)];
}
// put the block on the active list if it is not there yet
__device__ __forceinline__ void ensure_block(const Grid& g, int bx, int by, int bz) {
    int* t = &g.table[table_index(g, bx, by, bz)];
    int s = *(volatile int*)t;
    if (s != -1) return;
    int old = atomicCAS(t, -1, -2);
    if (old != -1) return;
    int slot = atomicAdd(g.n_slots, 1);
    if (slot >= g.cap) {
        g.flags[0] = 1;
        atomicExch(t, -1);
        return;
    }
    g.slot_coord[slot] = bx | (by << 10) | (bz << 20);
    __threadfence();
    atomicExch(t, slot);
}
// blocks covered by the 3^3 stencil of a particle at (x,y,z)
__device__ __forceinline__ void ensure_stencil_blocks(const Grid& g, float x, float y, float z) {
    int b0 = clampi(base_of(x, g.inv_dx), 0, g.n - 1), b1 = clampi(base_of(y, g.inv_dx), 0, g.n - 1),
        b2 = clampi(base_of(z, g.inv_dx), 0, g.n - 1);
    int x0 = b0 >> 2, x1 = min(b0 + 2, g.n - 1) >> 2;
    int y0 = b1 >> 2, y1 = min(b1 + 2, g.n - 1) >> 2;
    int z0 = b2 >> 2, z1 = min(b2 + 2, g.n - 1) >> 2;
    for (int a = x0; a <= x1; a++)
        for (int b = y0; b <= y1; b++)
            for (int c = z0; c <= z1; c++) ensure_block(g, a, b, c);
}
// per-axis terms of the node index: index(ix,iy,iz) = axis_offset<0>(ix) + axis_offset<1>(iy) + axis_offset<2>(iz)
template <int AXIS>
__device__ __forceinline__ int axis_offset(const Grid& g, int i) {
    const int blk = i >> 2, loc = i & 3;
    if (AXIS == 0) return blk * (g.nb * g.nb * BN) + (loc << 4);
    if (AXIS == 1) return blk * (g.nb * BN) + (loc << 2);
    return blk * BN + loc;
}
// index of node (ix,iy,iz) in the per-node arrays, -1 outside the grid
__device__ __forceinline__ int node_index(const Grid& g, int ix, int iy, int iz) {
    if ((unsigned)ix >= (unsigned)g.n || (unsigned)iy >= (unsigned)g.n || (unsigned)iz >= (unsigned)g.n) return -1;
    return table_index(g, ix >> 2, iy >> 2, iz >> 2) * BN + ((ix & 3) << 4) + ((iy & 3) << 2) + (iz & 3);
}
// node l (0..63: x = l>>4, y = (l>>2)&3, z = l&3) of the block with packed coordinates co
__device__ __forceinline__ int block_node(const Grid& g, int co, int l) {
    return table_index(g, co & 1023, (co >> 10) & 1023, (co >> 20) & 1023) * BN + l;
}
// ---- stencil addressing for a cell run.  The run's leader lane stores the nine per-axis offsets of the stencil whose
// packed cell is c (pack_cell: base + 2 per axis) -- o[0..2] x, o[3..5] y, o[6..8] z -- and every lane then adds the
// three that belong to its node (li,lj,lk) = (lane/9, lane/3%3, lane%3); lanes 27..31 alias node 26 (their loads hit
// the same shared-memory words as lane 26: no extra bank conflict; their results are discarded).  o[0] = -1: the
// stencil leaves the grid (nothing is loaded / scattered).
__device__ __forceinline__ void stencil_offsets(const Grid& g, int c, int* o) {
    const int bx = (c & 1023) - 2, by = ((c >> 10) & 1023) - 2, bz = (int)((unsigned)c >> 20) - 2;
    const bool ok = c >= 0 && (unsigned)bx <= (unsigned)(g.n - 3) && (unsigned)by <= (unsigned)(g.n - 3) && (unsigned)bz <= (unsigned)(g.n - 3);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o[i] = ok ? axis_offset<0>(g, bx + i) : -1;
        o[3 + i] = axis_offset<1>(g, by + i);
        o[6 + i] = axis_offset<2>(g, bz + i);
    }
}
struct LaneNode {
    int li, lj, lk;
    __device__ __forceinline__ explicit LaneNode(int lane) {
        const int l = min(lane, 26);
        li = l / 9; lj = 3 + (l / 3) % 3; lk = 6 + l % 3;
    }
    // node index from a run's nine offsets; negative if the run's stencil leaves the grid
    __device__ __forceinline__ int node(const int* o, bool& ok) const {
        const int x = o[li];
        ok = x >= 0;
        return x + o[lj] + o[lk];
    }
};
// activity of the (up to) 2x2x2 blocks under a 3^3 stencil whose base node is (bx,by,bz) >= 0, as an 8-bit mask
// (bit (cx<<2 | cy<<1 | cz)): eight independent table loads issued together (body / joint scatters, which must not
// touch inactive blocks)
__device__ __forceinline__ unsigned load_active8(const Grid& g, int bx, int by, int bz) {
    const int X0 = bx >> 2, Y0 = by >> 2, Z0 = bz >> 2;
    int sl[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int X = X0 + (c >> 2), Y = Y0 + ((c >> 1) & 1), Z = Z0 + (c & 1);
        bool ok = bx >= 0 && by >= 0 && bz >= 0 && X < g.nb && Y < g.nb && Z < g.nb;
        sl[c] = ok ? lookup_slot(g, X, Y, Z) : -1;
    }
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) m |= (sl[c] >= 0 ? 1u : 0u) << c;
    return m;
}
// per axis: does stencil offset i (0..2) from base b fall into the second block?  bit i of the result
__device__ __forceinline__ unsigned axis_cross(int b) {
    const int r = b & 3;  // offsets i with r + i >= 4
    return r == 3 ? 6u : (r == 2 ? 4u : 0u);
}

// quadratic B-spline factor of stencil offset i at fractional position f (mpm_utils.py:506-514):
// w = A (f - s)^2 + B, dw = 2A (f - s) with (A,s,B) = (.5,1.5,0), (-1,1,.75), (.5,.5,0)
__device__ __forceinline__ void bspline(float f, int i, float& w, float& dw) {
    float s = 1.5f - 0.5f * (float)i;
    float A = (i == 1) ? -1.0f : 0.5f;
    float B = (i == 1) ? 0.75f : 0.0f;
    float t = f - s;
    w = A * t * t + B;
    dw = 2.0f * A * t;
}

// ------------------------------------------------------------------ packed fp32 (FFMA2 / FMUL2 / FADD2)
// sm_100 issues two IEEE fp32 operations per instruction on an aligned register pair, with a scalar
// operand broadcast to both halves for free; the 27-node inner loops are written on float2 so that
// the issue slots per node halve.
__device__ __forceinline__ float2 fma2(float s, float2 a, float2 c) { return __ffma2_rn(make_float2(s, s), a, c); }
__device__ __forceinline__ float2 fma2(float2 s, float a, float2 c) { return __ffma2_rn(s, make_float2(a, a), c); }
__device__ __forceinline__ float2 mul2(float s, float2 a) { return __fmul2_rn(make_float2(s, s), a); }
__device__ __forceinline__ float2 mul2(float2 s, float a) { return __fmul2_rn(s, make_float2(a, a)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
struct V4 {
    float2 lo, hi;  // (x, y), (z, w)
};
__device__ __forceinline__ V4 v4(float x, float y, float z, float w) { return V4{make_float2(x, y), make_float2(z, w)}; }
__device__ __forceinline__ V4 mul4(float s, V4 a) { return V4{mul2(s, a.lo), mul2(s, a.hi)}; }
__device__ __forceinline__ V4 fma4(float s, V4 a, V4 c) { return V4{fma2(s, a.lo, c.lo), fma2(s, a.hi, c.hi)}; }
__device__ __forceinline__ V4 add4(V4 a, V4 b) { return V4{add2(a.lo, b.lo), add2(a.hi, b.hi)}; }

// ------------------------------------------------------------------ cell groups of a slab
// Particles are sorted by cell every few dozen substeps; in between they drift (a falling garment moves a cell or two
// between two re-sorts), and the particles of one OLD cell end up interleaved over two or more new cells.  The 32
// particles of a warp are therefore grouped by their CURRENT cell with one MATCH.ANY: every distinct cell is exactly one
// group ("run"), however its members are scattered over the lanes.  `cell` packs the clamped stencil base; lanes >= cnt
// must pass distinct negative values (they form trailing groups of their own).
struct Runs {
    unsigned starts;  // find_runs: bit l set = slot l is the first of a run.  group_runs: bit l set = lane l leads a group
    int nr;           // number of runs among the lanes < cnt
    int mine;         // run index of this lane
    int slot;         // group_runs: position of this lane's particle when the slab is ordered group by group
};
// adjacency version: the slab is already ordered run by run
__device__ __forceinline__ Runs find_runs(int lane, int cnt, int cell) {
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    Runs r;
    r.starts = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || cell != prev));
    r.nr = __popc(r.starts);
    r.mine = __popc(r.starts & (0xffffffffu >> (31 - lane))) - 1;
    r.slot = lane;
    return r;
}
// grouping version: runs are numbered by their leader (lowest) lane; slot = members of earlier groups + my rank in mine
__device__ __forceinline__ Runs group_runs(int lane, int cnt, int cell) {
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int leader = __ffs(peers) - 1;
    const unsigned lt = (1u << lane) - 1u;
    Runs r;
    r.starts = __ballot_sync(0xffffffffu, leader == lane && lane < cnt);
    r.nr = __popc(r.starts);
    r.mine = __popc(r.starts & ((1u << leader) - 1u));
    // exclusive prefix sum of the group sizes over the leader lanes, fetched from my leader
    int scan = (leader == lane) ? __popc(peers) : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, scan, o);
        if (lane >= o) scan += up;
    }
    const int excl = scan - ((leader == lane) ? __popc(peers) : 0);
    r.slot = __shfl_sync(0xffffffffu, excl, leader) + __popc(peers & lt);
    return r;
}
__device__ __forceinline__ int pack_cell(int bx, int by, int bz) {
    return clampi(bx + 2, 0, 1023) | (clampi(by + 2, 0, 1023) << 10) | (clampi(bz + 2, 0, 1023) << 20);
}

// ------------------------------------------------------------------ TMA slab staging
// 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) between global memory and a warp's
// shared-memory slab, completing on a per-warp mbarrier (loads) or a bulk group (stores).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// 16-byte cp.async (SASS LDGSTS): global -> shared without passing through registers
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the same with src_bytes (0 or 16) read from global and the rest of the 16 bytes zero-filled
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start
// while its predecessor drains; pdl_wait() blocks until the predecessor's memory is visible, pdl_trigger()
// lets the successor's CTAs be scheduled as soon as every CTA of this grid has started
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bytes of a slab of cnt records of F floats, rounded up to the 16-byte bulk-copy granule
// (arrays carry 32 records of slack, so over-reading / over-writing the tail is harmless)
__device__ __forceinline__ uint32_t slab_bytes(int cnt, int F) { return (uint32_t)((cnt * F * 4 + 15) & ~15); }

}  // namespace mpm
