// libmpm_b200.so -- host side of the B200 MPM substep solver and its C-ABI (include/mpm_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl.so.2 is dlopen'ed by mpm_attach_comm (no link-time dependency)

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpm_b200.h"
#include "mpm_kernels.cuh"

using namespace mpm;

static std::string g_create_error;

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            throw std::string(buf_);                                                                   \
        }                                                                                              \
    } while (0)

namespace {

struct Canon {  // canonical (reference-layout) device mirror
    float *x, *v, *C, *F, *Ft, *stress, *d, *Rinv, *faces, *vforce, *vol, *mass, *mu, *lam, *gamma, *kappa, *ys;
};

// ---------------------------------------------------------------- sort / import / export kernels
__global__ void k_keys(Grid g, int n, const float* __restrict__ x, int offset, uint32_t* keys, uint32_t* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = x + 3 * (size_t)(offset + i);
    keys[i] = sort_key(g, p[0], p[1], p[2]);
    vals[i] = i;
}
__global__ void k_invert(int n, const uint32_t* __restrict__ perm, int* __restrict__ inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = i;
}
// ---- direct re-sort: keys from the sorted records themselves, then one gather per class from copy A to copy B
__global__ void k_keys_rec(Grid g, int n, const float* __restrict__ rec, int F, uint32_t* keys, uint32_t* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rec + (size_t)i * F;
    keys[i] = sort_key(g, p[0], p[1], p[2]);
    vals[i] = i;
}
struct Recs;
struct Recs {  // sorted particle arrays (see mpm_device.cuh)
    // cloth elements: float4 / int4 streams
    int4* EFM;
    float4 *K0, *K1, *XE, *C0, *C1, *SP3, *EV, *ED1, *ED2;
    float4* D3[2];  // ping-pong d3 (+ C[8])
    int* CE;
    // traditional particles / vertices: sub-records
    float *TP, *TS, *TF, *VP;
    float4* VF[2];  // ping-pong vertex-force accumulators (buffer `cur` is filled by the substep that reads directions `cur`)
    int* CV;        // packed stencil base cell per vertex (lets G2P start its node loads early)
    // sorted slot -> canonical index within the class, and back
    uint32_t *permE, *permT, *permV;
    int *invE, *invT, *invV;
};
// ord[i] = old slot of the particle that moves to slot i
__global__ void k_permute_V(int Nv, const uint32_t* __restrict__ ord, Recs A, Recs B, int* __restrict__ o2n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nv) return;
    const int s = (int)ord[i];
    const float4* src = reinterpret_cast<const float4*>(A.VP + (size_t)s * VP_F);
    float4* dst = reinterpret_cast<float4*>(B.VP + (size_t)i * VP_F);
#pragma unroll
    for (int k = 0; k < VP_F / 4; k++) dst[k] = src[k];
    B.VF[0][i] = A.VF[0][s];
    B.VF[1][i] = A.VF[1][s];
    B.CV[i] = A.CV[s];
    const uint32_t c = A.permV[s];
    B.permV[i] = c;
    B.invV[c] = i;
    o2n[s] = i;
}
__global__ void k_permute_E(int Ne, const uint32_t* __restrict__ ord, Recs A, Recs B, const int* __restrict__ o2nV) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ne) return;
    const int s = (int)ord[i];
    int4 efm = A.EFM[s];  // corner slots follow the vertices to their new slots
    efm.x = o2nV[efm.x]; efm.y = o2nV[efm.y]; efm.z = o2nV[efm.z];
    B.EFM[i] = efm;
    B.K0[i] = A.K0[s]; B.K1[i] = A.K1[s]; B.XE[i] = A.XE[s]; B.EV[i] = A.EV[s]; B.ED1[i] = A.ED1[s]; B.ED2[i] = A.ED2[s];
    B.C0[i] = A.C0[s]; B.C1[i] = A.C1[s]; B.SP3[i] = A.SP3[s]; B.D3[0][i] = A.D3[0][s]; B.D3[1][i] = A.D3[1][s];
    B.CE[i] = A.CE[s];
    const uint32_t c = A.permE[s];
    B.permE[i] = c;
    B.invE[c] = i;
}
__global__ void k_permute_T(int Nt, const uint32_t* __restrict__ ord, Recs A, Recs B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nt) return;
    const int s = (int)ord[i];
    for (int k = 0; k < KP_F; k++) B.TP[(size_t)i * KP_F + k] = A.TP[(size_t)s * KP_F + k];
    for (int k = 0; k < S_F; k++) B.TS[(size_t)i * S_F + k] = A.TS[(size_t)s * S_F + k];
    for (int k = 0; k < TF_F; k++) B.TF[(size_t)i * TF_F + k] = A.TF[(size_t)s * TF_F + k];
    const uint32_t c = A.permT[s];
    B.permT[i] = c;
    B.invT[c] = i;
}
__global__ void k_import_E(Grid g, int Ne, const uint32_t* __restrict__ perm, Canon c, Recs R, int cur, const int* __restrict__ invV) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ne) return;
    int s = perm[i];  // canonical element index == canonical particle index
    const float x = c.x[3 * s], y = c.x[3 * s + 1], z = c.x[3 * s + 2];
    R.XE[i] = make_float4(x, y, z, 0.f);
    R.EV[i] = make_float4(c.v[3 * s], c.v[3 * s + 1], c.v[3 * s + 2], 0.f);
    const float* C = c.C + 9 * (size_t)s;
    R.C0[i] = make_float4(C[0], C[1], C[2], C[3]);
    R.C1[i] = make_float4(C[4], C[5], C[6], C[7]);
    const float* ds = c.d + 9 * (size_t)s;  // row-major 3x3 whose COLUMNS are d1,d2,d3
    R.ED1[i] = make_float4(ds[0], ds[3], ds[6], 0.f);
    R.ED2[i] = make_float4(ds[1], ds[4], ds[7], 0.f);
    R.D3[cur][i] = make_float4(ds[2], ds[5], ds[8], C[8]);
    R.SP3[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // int(face[k]), mpm_utils.py:172; the mass rides in the fourth word
    R.EFM[i] = make_int4(invV[(int)c.faces[3 * s]], invV[(int)c.faces[3 * s + 1]], invV[(int)c.faces[3 * s + 2]], __float_as_int(c.mass[s]));
    R.CE[i] = pack_cell(base_of(x, g.inv_dx), base_of(y, g.inv_dx), base_of(z, g.inv_dx));
    R.K0[i] = make_float4(c.Rinv[3 * s], c.Rinv[3 * s + 1], c.Rinv[3 * s + 2], c.mu[s]);
    R.K1[i] = make_float4(c.lam[s], c.gamma[s], c.kappa[s], c.vol[s]);
}
__global__ void k_import_T(int Nt, int Ne, const uint32_t* __restrict__ perm, Canon c, Recs R) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nt) return;
    int s = Ne + perm[i];
    float* p = R.TP + (size_t)i * KP_F;
    for (int k = 0; k < 3; k++) { p[P_X + k] = c.x[3 * s + k]; p[P_V + k] = c.v[3 * s + k]; }
    p[P_M] = c.mass[s];
    p[P_VOL] = c.vol[s];
    for (int k = 0; k < 9; k++) { p[P_C + k] = c.C[9 * (size_t)s + k]; R.TS[(size_t)i * S_F + k] = 0.f; }
    float* t = R.TF + (size_t)i * TF_F;
    for (int k = 0; k < 9; k++) { t[T_F + k] = c.F[9 * (size_t)s + k]; t[T_FT + k] = c.Ft[9 * (size_t)s + k]; }
    t[T_MU] = c.mu[s]; t[T_LAM] = c.lam[s]; t[T_YS] = c.ys[s];
}
__global__ void k_import_V(Grid g, int Nv, int Nnv, const uint32_t* __restrict__ perm, Canon c, Recs R) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nv) return;
    int s = Nnv + perm[i];
    float* p = R.VP + (size_t)i * VP_F;
    for (int k = 0; k < 3; k++) { p[V_X + k] = c.x[3 * s + k]; p[V_V + k] = c.v[3 * s + k]; }
    p[V_M] = c.mass[s];
    for (int k = 0; k < 9; k++) p[V_C + k] = c.C[9 * (size_t)s + k];
    R.VF[0][i] = R.VF[1][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    R.CV[i] = pack_cell(base_of(p[0], g.inv_dx), base_of(p[1], g.inv_dx), base_of(p[2], g.inv_dx));
}
// Element state in canonical order.  stepped = a substep has run since the import: the stress of the LAST substep is
// SP3 (x) the return-mapped d3 still held by direction buffer cur^1 (mpm_utils.py:177).
__global__ void k_export_E(int Ne, const uint32_t* __restrict__ perm, Canon c, Recs R, int cur, int stepped) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ne) return;
    int s = perm[i];
    const float4 xe = R.XE[i], ev = R.EV[i], e1 = R.ED1[i], e2 = R.ED2[i];
    const float x[3] = {xe.x, xe.y, xe.z}, v[3] = {ev.x, ev.y, ev.z}, d1[3] = {e1.x, e1.y, e1.z}, d2[3] = {e2.x, e2.y, e2.z};
    for (int k = 0; k < 3; k++) { c.x[3 * s + k] = x[k]; c.v[3 * s + k] = v[k]; }
    const float4 c0 = R.C0[i], c1 = R.C1[i], d3 = R.D3[cur][i];
    float* C = c.C + 9 * (size_t)s;
    C[0] = c0.x; C[1] = c0.y; C[2] = c0.z; C[3] = c0.w; C[4] = c1.x; C[5] = c1.y; C[6] = c1.z; C[7] = c1.w; C[8] = d3.w;
    float* dd = c.d + 9 * (size_t)s;
    for (int row = 0; row < 3; row++) { dd[3 * row] = d1[row]; dd[3 * row + 1] = d2[row]; }
    dd[2] = d3.x; dd[5] = d3.y; dd[8] = d3.z;
    if (stepped) {
        const float4 sp = R.SP3[i], n3 = R.D3[cur ^ 1][i];
        const float P3[3] = {sp.x, sp.y, sp.z}, nd3[3] = {n3.x, n3.y, n3.z};
        for (int rr = 0; rr < 3; rr++)
            for (int cc = 0; cc < 3; cc++) c.stress[9 * (size_t)s + 3 * rr + cc] = P3[rr] * nd3[cc];
    }
}
__global__ void k_export_T(int Nt, int Ne, const uint32_t* __restrict__ perm, Canon c, Recs R) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nt) return;
    int s = Ne + perm[i];
    const float* p = R.TP + (size_t)i * KP_F;
    const float* t = R.TF + (size_t)i * TF_F;
    for (int k = 0; k < 3; k++) { c.x[3 * s + k] = p[P_X + k]; c.v[3 * s + k] = p[P_V + k]; }
    for (int k = 0; k < 9; k++) {
        c.C[9 * (size_t)s + k] = p[P_C + k];
        c.stress[9 * (size_t)s + k] = R.TS[(size_t)i * S_F + k];
        c.F[9 * (size_t)s + k] = t[T_F + k];
        c.Ft[9 * (size_t)s + k] = t[T_FT + k];
    }
    c.mu[s] = t[T_MU]; c.lam[s] = t[T_LAM]; c.ys[s] = t[T_YS];  // damage / hardening mutate these
}
// vertex_force of the LAST substep (the reference zeroes it at the start of p2g2p, so it holds the last substep's cloth
// forces afterwards): buffer cur^1 once a substep has run since the import
__global__ void k_export_V(int Nv, int Nnv, const uint32_t* __restrict__ perm, Canon c, Recs R, int cur, int have_prev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nv) return;
    int s = Nnv + perm[i];
    int vl = perm[i];
    const float* p = R.VP + (size_t)i * VP_F;
    for (int k = 0; k < 3; k++) { c.x[3 * s + k] = p[V_X + k]; c.v[3 * s + k] = p[V_V + k]; }
    for (int k = 0; k < 9; k++) c.C[9 * (size_t)s + k] = p[V_C + k];
    const float4 f = have_prev ? R.VF[cur ^ 1][i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float fr[3] = {f.x, f.y, f.z};
    for (int k = 0; k < 3; k++) c.vforce[3 * vl + k] = fr[k];
}
__global__ void k_alloc_blocks(Grid g, int n, const float* __restrict__ rec, int F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = rec + (size_t)i * F;
    ensure_stencil_blocks(g, x[0], x[1], x[2]);
}
__global__ void k_fill_int(int* p, size_t n, int v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// dense debug export of the sparse grid
__global__ void k_export_grid(Grid g, float* gm, float* gvin, float* gvout) {
    const int n_slots = min(*g.n_slots, g.cap);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_slots * BN) return;
    int slot = idx >> 6, l = idx & 63, co = g.slot_coord[slot];
    int ix = ((co & 1023) << 2) + (l >> 4), iy = (((co >> 10) & 1023) << 2) + ((l >> 2) & 3), iz = (((co >> 20) & 1023) << 2) + (l & 3);
    if (ix >= g.n || iy >= g.n || iz >= g.n) return;
    size_t gi = ((size_t)ix * g.n + iy) * g.n + iz;
    idx = block_node(g, co, l);
    if (g.dbg_acc) {
        float4 a = g.dbg_acc[idx];
        if (gm) gm[gi] = a.w;
        if (gvin) { gvin[3 * gi] = a.x; gvin[3 * gi + 1] = a.y; gvin[3 * gi + 2] = a.z; }
    }
    if (gvout) { float4 v = g.vout[idx]; gvout[3 * gi] = v.x; gvout[3 * gi + 1] = v.y; gvout[3 * gi + 2] = v.z; }
}
// distinct nodes in the union of all particle stencils (SURVEY 8d "A")
__global__ void k_mark_nodes(Grid g, int n, const float* __restrict__ rec, int F, unsigned long long* mask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = rec + (size_t)i * F;
    int bx = base_of(x[0], g.inv_dx), by = base_of(x[1], g.inv_dx), bz = base_of(x[2], g.inv_dx);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            for (int c = 0; c < 3; c++) {
                int ni = node_index(g, bx + a, by + b, bz + c);
                if (ni >= 0) atomicOr(&mask[ni >> 6], 1ull << (ni & 63));
            }
}
__global__ void k_popc(const unsigned long long* mask, int n, unsigned long long* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = (i < n) ? __popcll(mask[i]) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// blocks under the stencil of any particle displaced by up to `margin` cells in every direction: what this
// rank can activate before the shared-block list is rebuilt
// g_bit: the value written (1, or 1 << rank so that a byte-sum over <= 8 ranks is the set of ranks that can touch a block)
// markj (optional): blocks reachable by prescribed-velocity (joint) particles, the first njoint of the class in
// canonical order -- only those blocks carry mover accumulators
__global__ void k_mark_potential(Grid g, int n, const float* __restrict__ rec, int F, int margin, unsigned char* __restrict__ mark,
                                 unsigned char* __restrict__ markj = nullptr, const uint32_t* __restrict__ perm = nullptr, int njoint = 0,
                                 int g_bit = 1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool joint = markj && perm && (int)perm[i] < njoint;
    const unsigned char bit = (unsigned char)g_bit;
    const float* x = rec + (size_t)i * F;
    int lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        int b = clampi(base_of(x[a], g.inv_dx), 0, g.n - 1);
        lo[a] = clampi(b - margin, 0, g.n - 1) >> 2;
        hi[a] = clampi(b + 2 + margin, 0, g.n - 1) >> 2;
    }
    for (int a = lo[0]; a <= hi[0]; a++)
        for (int b = lo[1]; b <= hi[1]; b++)
            for (int c = lo[2]; c <= hi[2]; c++) {
                mark[table_index(g, a, b, c)] = bit;
                if (joint) markj[table_index(g, a, b, c)] = bit;
            }
}
// ---- sharded runs: the grid blocks shared with other ranks travel through one packed buffer
// [n_shared][64 nodes][acc float4 | mov float4]; inactive blocks pack zeros and ignore the result
// n_dev (device int) overrides n_shared when non-null: the captured sharded graphs keep a fixed launch geometry
// (capacity) while the shared list is rebuilt underneath them
__global__ void k_shared_pack(Grid g, const int* __restrict__ shared, int n_shared, const int* __restrict__ n_dev, float4* __restrict__ buf) {
    if (n_dev) n_shared = *n_dev;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_shared * BN; idx += gridDim.x * blockDim.x) {
        const int co = shared[idx >> 6], l = idx & 63;
        const int blk = table_index(g, co & 1023, (co >> 10) & 1023, (co >> 20) & 1023);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), m = a;
        const int ni = block_node(g, co, l);
        if (g.table[blk] >= 0) { a = g.acc[ni]; m = g.mov[ni]; }
        buf[2 * idx] = a;
        buf[2 * idx + 1] = m;
    }
}
__global__ void k_shared_unpack(Grid g, const int* __restrict__ shared, int n_shared, const int* __restrict__ n_dev, const float4* __restrict__ buf) {
    if (n_dev) n_shared = *n_dev;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_shared * BN; idx += gridDim.x * blockDim.x) {
        const int co = shared[idx >> 6], l = idx & 63;
        const int blk = table_index(g, co & 1023, (co >> 10) & 1023, (co >> 20) & 1023);
        const int ni = block_node(g, co, l);
        if (g.table[blk] >= 0) { g.acc[ni] = buf[2 * idx]; g.mov[ni] = buf[2 * idx + 1]; }
    }
}

// device-side rebuild of the shared-block list: mark[] holds, after a byte-wise sum over the ranks, how many ranks
// can touch each block; the blocks with count >= 2 are compacted in ascending order (the same list on every rank)
struct SharedPred {
    const unsigned char* mark;
    const unsigned char* markj;  // non-null: additionally some rank has joint particles there
    int bits;                    // 1: mark[] is a set of rank bits, 0: a count
    __device__ bool operator()(int i) const {
        const int m = mark[i];
        return (bits ? __popc(m) >= 2 : m >= 2) && (!markj || markj[i] >= 1);
    }
};
// in-graph exchange buffer: [capA blocks x 64 acc float4 | capM blocks x 64 mov float4]; list A = blocks at least
// two ranks can touch, list M = the subset that can carry mover (joint) accumulators
struct SharedLists {
    const int *A, *nA, *M, *nM;
    int capA, capM;
};
__global__ void k_shared_pack2(Grid g, SharedLists L, float4* __restrict__ buf) {
    const int nA = min(*L.nA, L.capA) * BN, nM = min(*L.nM, L.capM) * BN;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nA + nM; idx += gridDim.x * blockDim.x) {
        const bool mv = idx >= nA;
        const int j = mv ? idx - nA : idx;
        const int co = (mv ? L.M : L.A)[j >> 6], l = j & 63;
        const int blk = table_index(g, co & 1023, (co >> 10) & 1023, (co >> 20) & 1023);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ni = block_node(g, co, l);
        if (g.table[blk] >= 0) a = (mv ? g.mov : g.acc)[ni];
        buf[mv ? (size_t)L.capA * BN + j : j] = a;
    }
}
__global__ void k_shared_unpack2(Grid g, SharedLists L, const float4* __restrict__ buf) {
    const int nA = min(*L.nA, L.capA) * BN, nM = min(*L.nM, L.capM) * BN;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nA + nM; idx += gridDim.x * blockDim.x) {
        const bool mv = idx >= nA;
        const int j = mv ? idx - nA : idx;
        const int co = (mv ? L.M : L.A)[j >> 6], l = j & 63;
        const int blk = table_index(g, co & 1023, (co >> 10) & 1023, (co >> 20) & 1023);
        const int ni = block_node(g, co, l);
        if (g.table[blk] >= 0) (mv ? g.mov : g.acc)[ni] = buf[mv ? (size_t)L.capA * BN + j : j];
    }
}
__global__ void k_shared_coords(Grid g, const int* __restrict__ lin, int* __restrict__ n_sel, int cap, int* __restrict__ coords,
                                const unsigned char* __restrict__ mark, unsigned char* __restrict__ members, int* __restrict__ map) {
    const int n = *n_sel;
    if (n > cap && blockIdx.x == 0 && threadIdx.x == 0) g.flags[0] = 1;  // reported as overflow
    const int m = min(n, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int t = lin[i];
        const int bz = t % g.nb, by = (t / g.nb) % g.nb, bx = t / (g.nb * g.nb);
        coords[i] = bx | (by << 10) | (bz << 20);
        members[i] = mark[t];  // the ranks that can touch the block (peer-to-peer exchange)
        map[t] = i;            // block -> position in this list (the grid update's fused pull), -1 elsewhere
    }
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

struct MpmSolver {
    MpmConfig cfg{};
    int N = 0, Ne = 0, Nt = 0, Nv = 0, Nnv = 0;
    Grid g{};
    ModelDev md{};
    std::string err;
    // particles
    Recs R{};
    Recs R2{};  // the other copy: a re-sort gathers R -> R2 and swaps them
    int rbuf = 0;  // which copy R is (the captured graphs hold raw pointers: part of their keys)
    uint32_t *ordE = nullptr, *ordT = nullptr, *ordV = nullptr;  // re-sort: old slot of every new slot
    int* o2nV = nullptr;                                         // re-sort: new slot of every old vertex slot
    uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    Canon canon{};
    std::vector<void*> allocs;
    // body mesh / joints
    int* mesh_faces = nullptr;
    float *mesh_x = nullptr, *mesh_v = nullptr;
    float *joint_t = nullptr, *joint_v = nullptr, *joint_f = nullptr;
    bool has_collider = false, has_mover = false;
    float col_friction = 0.f;
    // BCs / ops
    BCDesc* d_bcs = nullptr;
    std::vector<BCDesc> h_bcs;
    ParticleOp* d_ops = nullptr;
    std::vector<ParticleOp> h_ops;
    bool bcs_dirty = false, ops_dirty = false;
    StepState* st = nullptr;
    // state flags
    bool have_state = false, need_sort = true, canon_stale = false;
    int since_sort = 0, resort_interval = 64;
    unsigned char* d_mark = nullptr;  // sharded runs: potential-block marks
    std::vector<unsigned char> h_mark;
    int* d_shared = nullptr;  // sharded runs: coordinates of the blocks shared with other ranks
    int n_shared = 0, shared_cap = 0, shared_age = -1;
    int* d_n_shared = nullptr;     // device copy of n_shared (read by the captured pack / unpack launches)
    ncclComm_t comm = nullptr;     // mpm_attach_comm: this solver's own communicator
    MpmHostAllGatherFn host_ag = nullptr;  // mpm_attach_host_comm: the caller's blocking host all-gather (gloo, MPI, ...)
    void* host_ctx = nullptr;
    std::vector<unsigned char> host_stage;
    unsigned char* ag_dev = nullptr;  // device staging of the NCCL all-gather of host bytes
    size_t ag_bytes = 0;
    int comm_rank = 0, comm_size = 1;
    float* xbuf = nullptr;         // exchange buffer of the in-graph path, xcap_blocks * 512 floats
    int xcap_blocks = 0;
    void* shard_graphs = nullptr;  // cache of captured sharded windows
    int* d_sel = nullptr;          // compaction output (linear block indices), [nb^3]
    void* sel_tmp = nullptr;
    size_t sel_bytes = 0;
    int* h_nshared = nullptr;      // pinned mirror of the device-side counts {A, M} (read one rebuild late)
    int *d_sharedM = nullptr, *d_nM = nullptr;  // list M (mover blocks)
    int xcapM = 0;
    unsigned char *d_memA = nullptr, *d_memM = nullptr;  // member ranks of every listed block
    bool use_p2p = true, p2p_ready = false;  // MPM_B200_P2P=0: ncclAllReduce instead of the peer-to-peer exchange
    unsigned char* peer_local = nullptr;     // this rank's receive area (cudaMalloc, exported over CUDA IPC)
    void* peer_open[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    PeerArea peer{};
    PeerArea gu_peer{};  // what the grid update of the substep being launched pulls from (nranks = 0: nothing)
    int *d_mapA = nullptr, *d_mapM = nullptr;  // block -> position in the shared lists
    int cur = 0;             // direction (D3) / vertex-force (VF) buffer of the coming substep
    bool have_prev = false;  // a substep has run since the last import: buffer cur^1 holds the return-mapped d3 and the vertex
                             // forces of the last substep
    int n_resorts = 0, n_rebuilds = 0;
    long long n_substeps = 0;
    int launches = 0;
    double host_time = 0.0;
    bool debug = false, profiling = false;
    bool pending_mover = false;  // mover flag of the scatter half, consumed by the gather half
    int pending_njt = 0;
    unsigned long long* node_mask = nullptr;
    // device block count mirrored (asynchronously) into pinned host memory after each re-sort
    int* h_nslots = nullptr;
    int slots_seen = 0;
    void* graph_cache_ptr = nullptr;
    bool use_graphs = true;
    bool use_pdl = true;  // MPM_B200_PDL=0 disables programmatic dependent launch
    bool scatter_early = false;  // MPM_B200_SCATTER_EARLY=1
    float scatter_at = 0.3f;     // MPM_B200_SCATTER_AT: position of the body-scatter CTAs in the vertex P2G grid (fraction; < 0: own launch).
                                 // C3, us per substep: own launch 78.7; at 0.0 / 0.3 / 0.5 / 0.7 / 0.9 of the grid 76.5 / 75.9 / 76.6 / 77.5 / 77.3
    cudaStream_t cap_stream = nullptr;
    // profiling
    cudaEvent_t ev[10]{};
    MpmProfile prof{};
    std::vector<std::array<cudaEvent_t, 9>> pending;

    template <typename T>
    T* dalloc(size_t n) {
        void* p = nullptr;
        CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        CK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
        allocs.push_back(p);
        return (T*)p;
    }
};

// ---------------------------------------------------------------- internals
static int sort_bits(MpmSolver* s) {
    int bits = 6;
    for (int nb = s->g.nb; nb > 1; nb = (nb + 1) / 2) bits += 3;
    return std::min(bits + 3, 32);
}
static void sort_class(MpmSolver* s, int n, int offset, uint32_t* perm, int* inv, cudaStream_t q) {
    if (n == 0) return;
    k_keys<<<cdiv(n, 256), 256, 0, q>>>(s->g, n, s->canon.x, offset, s->keys_in, s->vals_in);
    size_t tmp = s->cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(s->cub_tmp, tmp, s->keys_in, s->keys_out, s->vals_in, perm, n, 0, sort_bits(s), q));
    k_invert<<<cdiv(n, 256), 256, 0, q>>>(n, perm, inv);
    s->launches += 3;
}

static void export_to_canon(MpmSolver* s, cudaStream_t q) {
    if (!s->have_state || !s->canon_stale) return;
    if (s->Ne) k_export_E<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->Ne, s->R.permE, s->canon, s->R, s->cur, s->have_prev ? 1 : 0);
    if (s->Nt) k_export_T<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->Nt, s->Ne, s->R.permT, s->canon, s->R);
    if (s->Nv) k_export_V<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->Nv, s->Nnv, s->R.permV, s->canon, s->R, s->cur, s->have_prev ? 1 : 0);
    s->launches += 3;
    s->canon_stale = false;
}

// order of one class by the cell keys of its CURRENT sorted records: ord[i] = old slot of new slot i
static void sort_records(MpmSolver* s, int n, const float* rec, int F, uint32_t* ord, cudaStream_t q) {
    if (n == 0) return;
    k_keys_rec<<<cdiv(n, 256), 256, 0, q>>>(s->g, n, rec, F, s->keys_in, s->vals_in);
    size_t tmp = s->cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(s->cub_tmp, tmp, s->keys_in, s->keys_out, s->vals_in, ord, n, 0, sort_bits(s), q));
    s->launches += 2;
}

// Re-sort.  After an import the canonical mirror is the source (sort, then import into the records); the periodic
// re-sort of a running simulation never touches the mirror: the records are gathered straight from copy R into copy
// R2 in the new order (one read and one write of the state instead of export -> mirror -> import), the elements'
// corner slots and the slot <-> canonical maps are composed on the way, and the copies swap.  Both end with the rebuild
// of the sparse grid.
static void resort(MpmSolver* s, cudaStream_t q) {
    if (!s->need_sort && s->have_state && getenv("MPM_B200_RESORT_VIA_MIRROR") == nullptr) {
        const Recs &A = s->R, &B = s->R2;
        sort_records(s, s->Nv, A.VP, VP_F, s->ordV, q);
        if (s->Nv) k_permute_V<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->Nv, s->ordV, A, B, s->o2nV);
        sort_records(s, s->Ne, (const float*)A.XE, 4, s->ordE, q);
        if (s->Ne) k_permute_E<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->Ne, s->ordE, A, B, s->o2nV);
        sort_records(s, s->Nt, A.TP, KP_F, s->ordT, q);
        if (s->Nt) k_permute_T<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->Nt, s->ordT, A, B);
        s->launches += 3;
        std::swap(s->R, s->R2);
        s->rbuf ^= 1;
    } else {
    export_to_canon(s, q);
    sort_class(s, s->Ne, 0, s->R.permE, s->R.invE, q);
    sort_class(s, s->Nt, s->Ne, s->R.permT, s->R.invT, q);
    sort_class(s, s->Nv, s->Nnv, s->R.permV, s->R.invV, q);
    if (s->Ne) k_import_E<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->g, s->Ne, s->R.permE, s->canon, s->R, s->cur, s->R.invV);
    s->have_prev = false;
    if (s->Nt) k_import_T<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->Nt, s->Ne, s->R.permT, s->canon, s->R);
    if (s->Nv) k_import_V<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->g, s->Nv, s->Nnv, s->R.permV, s->canon, s->R);
    }
    // all accumulators are zero between substeps, so rebuilding the table needs no pool sweep
    size_t nt = (size_t)s->g.nb * s->g.nb * s->g.nb;
    k_fill_int<<<std::min(cdiv((long long)nt, 256), 1184), 256, 0, q>>>(s->g.table, nt, -1);
    CK(cudaMemsetAsync(s->g.n_slots, 0, sizeof(int), q));
    if (s->Ne) k_alloc_blocks<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->g, s->Ne, (const float*)s->R.XE, 4);
    if (s->Nt) k_alloc_blocks<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->g, s->Nt, s->R.TP, KP_F);
    if (s->Nv) k_alloc_blocks<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->g, s->Nv, s->R.VP, VP_F);
    CK(cudaMemcpyAsync(s->h_nslots, s->g.n_slots, sizeof(int), cudaMemcpyDeviceToHost, q));
    s->launches += 7;
    s->need_sort = false;
    s->since_sort = 0;
    s->n_resorts++;
    CK(cudaGetLastError());
}

static void upload_lists(MpmSolver* s, cudaStream_t q) {
    if (s->bcs_dirty) {
        if (!s->h_bcs.empty())
            CK(cudaMemcpyAsync(s->d_bcs, s->h_bcs.data(), s->h_bcs.size() * sizeof(BCDesc), cudaMemcpyHostToDevice, q));
        s->bcs_dirty = false;
    }
    if (s->ops_dirty) {
        if (!s->h_ops.empty())
            CK(cudaMemcpyAsync(s->d_ops, s->h_ops.data(), s->h_ops.size() * sizeof(ParticleOp), cudaMemcpyHostToDevice, q));
        s->ops_dirty = false;
    }
}

struct SubstepArgs {
    float dt;
    bool collider, mover, advance_mesh;
    int njt;
};

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor in
// the stream drains; it calls griddepcontrol.wait before touching global memory (mpm_device.cuh pdl_wait).
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t q, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = q;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

// the grid update grid-strides over the allocated blocks (count on the device): 8 CTAs of 256 threads per SM
constexpr int GRID_UPDATE_CTAS = 148 * 5;  // one wave at 48 registers x 256 threads
constexpr int GRID_UPDATE_CTAS_PEER = 148 * 4;  // k_grid_update<true>: every CTA resident at once (launch bounds 256 x 4)

// A substep has two halves around the grid: everything that SCATTERS to it (constitutive update, P2G,
// body-mesh and joint scatters) and everything that reads it back (grid update, G2P).  Single-GPU
// stepping runs them back to back; a sharded run reduces the shared grid blocks in between (mpm_shared_*).
enum { HALF_SCATTER = 1, HALF_GATHER = 2, HALF_BOTH = 3 };

static void launch_substep(MpmSolver* s, const SubstepArgs& a, cudaStream_t q, int halves = HALF_BOTH) {
    const int n_bc = (int)s->h_bcs.size(), n_ops = (int)s->h_ops.size();
    cudaEvent_t* ev = nullptr;
    std::array<cudaEvent_t, 9> evs;
    if (s->profiling && halves == HALF_BOTH) {
        for (auto& e : evs) CK(cudaEventCreate(&e));
        ev = evs.data();
        CK(cudaEventRecord(ev[0], q));
    }
    const Recs& R = s->R;
    auto sm = [](int nw, int wb) { return (size_t)(128 + nw * wb); };
    const int cur = s->cur;
    const bool pdl = s->use_pdl && !s->profiling;
    if (halves & HALF_SCATTER) {
    if (n_ops) {
        if (s->Ne) {
            k_particle_ops<<<cdiv(s->Ne, 256), 256, 0, q>>>(s->Ne, (const float*)R.XE, 4, (float*)R.EV, 4, (const float*)R.EFM + 3, 4, s->R.permE, 0, s->d_ops, n_ops, s->st, a.dt);
            s->launches++;
        }
        if (s->Nt) k_particle_ops<<<cdiv(s->Nt, 256), 256, 0, q>>>(s->Nt, R.TP + P_X, KP_F, R.TP + P_V, KP_F, R.TP + P_M, KP_F, s->R.permT, s->Ne, s->d_ops, n_ops, s->st, a.dt);
        if (s->Nv) k_particle_ops<<<cdiv(s->Nv, 256), 256, 0, q>>>(s->Nv, R.VP + V_X, VP_F, R.VP + V_V, VP_F, R.VP + V_M, VP_F, s->R.permV, s->Nnv, s->d_ops, n_ops, s->st, a.dt);
        s->launches += 2;
    }
    if (s->Nt) {
        k_stress_traditional<<<cdiv(s->Nt, 32 * STRESS_T_NW), 32 * STRESS_T_NW, sm(STRESS_T_NW, STRESS_T_WB), q>>>(s->Nt, R.TF, R.TS, s->md, a.dt);
        s->launches++;
    }
    if (ev) CK(cudaEventRecord(ev[1], q));
    const int ppb = 32 * P2G_NW;  // particles per P2G block
    if (s->Ne) {  // cloth stress + scatter; VF must be complete before the vertex scatter
        ElemIO io{R.EFM, R.K0, R.K1, R.XE, R.EV, R.ED1, R.ED2, R.C0, R.C1, R.D3[cur], R.SP3, R.VF[cur]};
        launch_pdl(k_p2g_elements, cdiv(s->Ne, ppb), ppb, P2G_SMEM, q, pdl, s->g, io, s->Ne, a.dt, s->md.rpic, s->md.friction_coeff);
        s->launches++;
    }
    if (s->Nt) {
        P2GIn in{R.TP, R.TS};
        launch_pdl(k_p2g<1>, cdiv(s->Nt, ppb), ppb, P2G_SMEM, q, pdl, s->g, in, s->Nt, a.dt, s->md.rpic, BodyScatter{});
        s->launches++;
    }
    // body-mesh collider and joint movers (mpm_solver.py:382-472): one launch behind the P2G kernels
    // (profiling: two launches, so that the two phases are timed separately)
    const int tot = a.mover ? a.njt + s->cfg.num_joint_v + s->cfg.num_joint_f : 0;
    const ColliderArgs ca{a.collider ? s->cfg.n_mesh_f : 0, s->mesh_faces, s->mesh_x, s->mesh_v, s->st, a.dt, a.advance_mesh ? 1 : 0};
    const MoverArgs ma{a.mover ? a.njt : 0, a.mover ? s->cfg.num_joint_v : 0, a.mover ? s->cfg.num_joint_f : 0, s->Nt,
                       s->joint_t, s->joint_v, s->joint_f, R.TP, R.VP, R.XE, s->R.invE, s->R.invT, s->R.invV};
    auto scatter_launch = [&]() {
        if (ev) {  // profiling: separate launches so that the two phases are timed separately
            if (ca.Mf) { k_collider_scatter<<<cdiv(ca.Mf, 128), 128, 0, q>>>(s->g, ca); s->launches++; }
            CK(cudaEventRecord(ev[3], q));
            if (tot) { k_mover_scatter<<<cdiv(tot, 128), 128, 0, q>>>(s->g, ma); s->launches++; }
            CK(cudaEventRecord(ev[4], q));
        } else if (ca.Mf + tot) {
            launch_pdl(k_body_scatter, cdiv(3LL * (ca.Mf + tot), 128), 128, 0, q, pdl, s->g, ca, ma);
            s->launches++;
        }
    };
    // where the body scatter runs: as CTAs in the middle of the vertex P2G grid (default), or as its own launch in front of
    // (MPM_B200_SCATTER_EARLY=1) / behind (MPM_B200_SCATTER_AT < 0, profiling, no cloth vertices) that kernel
    const bool fused = s->Nv && !ev && !s->scatter_early && s->scatter_at >= 0.0f && ca.Mf + tot > 0;
    const bool scatter_early = s->scatter_early && !ev;
    if (scatter_early) scatter_launch();
    if (s->Nv) {
        P2GIn in{R.VP, (const float*)R.VF[cur]};
        const int slabs = cdiv(s->Nv, 32 * P2G_V_NW);
        BodyScatter bs{};
        if (fused) {
            bs.ca = ca; bs.ma = ma;
            bs.n_ctas = cdiv(3LL * (ca.Mf + tot), 32 * P2G_V_NW);
            bs.first = std::min(slabs, (int)(s->scatter_at * slabs));
        }
        launch_pdl(k_p2g<2>, slabs + bs.n_ctas, 32 * P2G_V_NW, 128 + P2G_V_NW * P2G_WB, q, pdl, s->g, in, s->Nv, a.dt, s->md.rpic, bs);
        s->launches++;
    }
    if (ev) CK(cudaEventRecord(ev[2], q));
    if (!scatter_early && !fused) scatter_launch();
    }  // HALF_SCATTER
    if (!(halves & HALF_GATHER)) return;
    // one thread per node of the active blocks, grid-strided over the device-side block count
    if (s->gu_peer.nranks > 1)  // sharded, peer-to-peer: the exchange is fused in; all CTAs must be resident (4 per SM)
        launch_pdl(k_grid_update<true>, GRID_UPDATE_CTAS_PEER, 256, 0, q, pdl, s->g, s->md, a.dt, a.collider ? 1 : 0, s->col_friction, a.mover ? 1 : 0, (const BCDesc*)s->d_bcs, n_bc, (const StepState*)s->st, s->gu_peer);
    else
        launch_pdl(k_grid_update<false>, GRID_UPDATE_CTAS, 256, 0, q, pdl, s->g, s->md, a.dt, a.collider ? 1 : 0, s->col_friction, a.mover ? 1 : 0, (const BCDesc*)s->d_bcs, n_bc, (const StepState*)s->st, PeerArea{});
    s->launches++;
    if (ev) CK(cudaEventRecord(ev[5], q));
    // gather side: the first kernel waits for the grid update, the others follow it; the last kernel advances time
    Advance none{nullptr, nullptr, 0}, adv{s->st, s->d_bcs, n_bc};
    const int gpb = 32 * G2P_NW;
    int follower = 0;
    if (s->Nv) {
        launch_pdl(k_g2p_vertices, cdiv(s->Nv, gpb), gpb, sm(G2P_NW, G2P_V_WB), q, pdl, s->g, s->Nv, R.VP, R.VF[cur ^ 1], R.CV, a.dt, 1, (s->Nt || s->Ne) ? none : adv);
        s->launches++;
        follower = 1;
    }
    if (s->Nt) {
        launch_pdl(k_g2p_traditional, cdiv(s->Nt, gpb), gpb, sm(G2P_NW, G2P_T_WB), q, pdl, s->g, s->Nt, R.TP, R.TF, a.dt, (follower && pdl) ? 0 : 1, s->Ne ? none : adv);
        s->launches++;
    }
    if (ev) CK(cudaEventRecord(ev[6], q));
    if (s->Ne) {
        ElemG2P eg{R.EFM, R.XE, R.EV, R.ED1, R.ED2, R.C0, R.C1, R.D3[cur], R.D3[cur ^ 1], R.CE, R.VP};
        launch_pdl(k_g2p_elements, cdiv(s->Ne, gpb), gpb, sm(G2P_NW, G2P_E_WB), q, pdl, s->g, s->Ne, eg, a.dt, adv);
        s->launches++;
    }
    if (s->Ne) s->cur ^= 1;
    s->have_prev = true;
    if (ev) {
        CK(cudaEventRecord(ev[7], q));
        CK(cudaEventRecord(ev[8], q));
        s->pending.push_back(evs);
    }
}

// ---------------------------------------------------------------- CUDA graphs
// A captured run of GRAPH_U substeps is replayed instead of ~8 launches per substep; every
// per-substep quantity (time, substep index, block count) lives in device memory, so one
// instantiated graph serves every call with the same launch geometry.
constexpr int GRAPH_U = 16;  // even: the direction ping-pong index is the same before and after a replay
struct GraphKey {
    float dt;
    int collider, mover, advance_mesh, njt, cur, n_bc, n_ops, debug, rbuf, len;
    bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;
    int launches_per_replay;
};
static std::vector<GraphEntry>& graph_cache(MpmSolver* s);

static void drain_profile(MpmSolver* s) {
    for (auto& evs : s->pending) {
        CK(cudaEventSynchronize(evs[8]));
        float t[8];
        for (int i = 0; i < 8; i++) CK(cudaEventElapsedTime(&t[i], evs[i], evs[i + 1]));
        s->prof.stress_ms += t[0];
        s->prof.p2g_ms += t[1];
        s->prof.collider_scatter_ms += t[2];
        s->prof.mover_scatter_ms += t[3];
        s->prof.grid_ms += t[4];
        s->prof.g2p_v_ms += t[5];
        s->prof.g2p_e_ms += t[6];
        s->prof.n_substeps++;
        for (auto& e : evs) cudaEventDestroy(e);
    }
    s->pending.clear();
}

static std::vector<GraphEntry>& graph_cache(MpmSolver* s) {
    if (!s->graph_cache_ptr) s->graph_cache_ptr = new std::vector<GraphEntry>();
    return *static_cast<std::vector<GraphEntry>*>(s->graph_cache_ptr);
}
static void destroy_graphs(MpmSolver* s) {
    if (!s->graph_cache_ptr) return;
    auto& v = graph_cache(s);
    for (auto& e : v) cudaGraphExecDestroy(e.exec);
    delete &v;
    s->graph_cache_ptr = nullptr;
}
static void destroy_sharded(MpmSolver* s);  // communicator + captured sharded windows (defined with the NCCL path)
static void destroy_shard_graphs(MpmSolver* s);
// single = the call is ONE substep (the unchanged callers' p2g2p loop): it replays a one-substep graph
static void run_substeps(MpmSolver* s, SubstepArgs a, int count, cudaStream_t q, bool single = false) {
    const bool graphs = s->use_graphs && !s->profiling;
    while (count > 0) {
        const int len = (graphs && count >= GRAPH_U) ? GRAPH_U : ((graphs && single && count == 1 && !s->debug) ? 1 : 0);
        if (len) {
            GraphKey key{};
            key.dt = a.dt; key.collider = a.collider; key.mover = a.mover; key.advance_mesh = a.advance_mesh;
            key.njt = a.njt; key.cur = s->cur; key.n_bc = (int)s->h_bcs.size();
            key.n_ops = (int)s->h_ops.size(); key.debug = s->debug; key.rbuf = s->rbuf; key.len = len;
            auto& cache = graph_cache(s);
            GraphEntry* hit = nullptr;
            for (auto& e : cache) if (e.key == key) hit = &e;
            if (!hit) {
                cudaGraph_t graph;
                int before = s->launches;
                // capture on a private stream (the caller's stream may be the legacy default stream,
                // which cannot be captured); the instantiated graph is launched on the caller's stream
                if (!s->cap_stream) CK(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
                CK(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
                const int cur0 = s->cur;
                const bool prev0 = s->have_prev;
                for (int i = 0; i < len; i++) launch_substep(s, a, s->cap_stream);
                s->cur = cur0;  // the capture only recorded launches: the replay below does the stepping
                s->have_prev = prev0;
                CK(cudaStreamEndCapture(s->cap_stream, &graph));
                GraphEntry e;
                e.key = key;
                e.launches_per_replay = s->launches - before;
                s->launches = before;
                CK(cudaGraphInstantiate(&e.exec, graph, 0));
                cudaGraphDestroy(graph);
                if (cache.size() > 32) { for (auto& c : cache) cudaGraphExecDestroy(c.exec); cache.clear(); }
                cache.push_back(e);
                hit = &cache.back();
            }
            CK(cudaGraphLaunch(hit->exec, q));
            s->have_prev = true;
            if (s->Ne && (len & 1)) s->cur ^= 1;
            s->launches += hit->launches_per_replay;
            count -= len;
        } else {
            launch_substep(s, a, q);
            count -= 1;
        }
    }
}

// ---------------------------------------------------------------- C-ABI
#define API_BEGIN(s)              \
    if (!(s)) return -1;          \
    try {                         \
        CK(cudaSetDevice((s)->cfg.device));
#define API_END(s)                \
    }                             \
    catch (const std::string& e) { \
        (s)->err = e;             \
        return -2;                \
    }                             \
    return 0;

extern "C" {

int mpm_create(const MpmConfig* cfg, MpmSolver** out) {
    if (!cfg || !out) return -1;
    MpmSolver* s = new MpmSolver();
    try {
        s->cfg = *cfg;
        CK(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major < 9) throw std::string("mpm_b200 needs sm_90+ vector atomics; built for sm_100a");
        s->N = cfg->n_particles; s->Ne = cfg->n_elements; s->Nv = cfg->n_vertices;
        s->Nnv = s->N - s->Nv; s->Nt = s->Nnv - s->Ne;
        if (s->N <= 0 || s->Ne < 0 || s->Nv < 0 || s->Nt < 0 || cfg->n_grid < 8 || cfg->n_grid > 1020)
            throw std::string("invalid particle counts or n_grid (8..1020)");
        if (cfg->resort_interval > 0) s->resort_interval = cfg->resort_interval;
        if (const char* e = getenv("MPM_B200_RESORT")) { if (cfg->resort_interval <= 0 && atoi(e) > 0) s->resort_interval = atoi(e); }
        if (const char* e = getenv("MPM_B200_PDL")) s->use_pdl = atoi(e) != 0;
        if (const char* e = getenv("MPM_B200_SCATTER_EARLY")) s->scatter_early = atoi(e) != 0;
        if (const char* e = getenv("MPM_B200_SCATTER_AT")) s->scatter_at = (float)atof(e);
        if (const char* e = getenv("MPM_B200_GRAPHS")) s->use_graphs = atoi(e) != 0;
        if (const char* e = getenv("MPM_B200_P2P")) s->use_p2p = atoi(e) != 0;
        Grid& g = s->g;
        g.n = cfg->n_grid;
        g.nb = (g.n + BS - 1) / BS;
        g.lim = cfg->grid_lim;
        // dx, inv_dx exactly as init_other_params: python doubles rounded to f32 (mpm_data_structure.py:692-697)
        g.dx = (float)((double)cfg->grid_lim / (double)cfg->n_grid);
        g.inv_dx = (float)((double)cfg->n_grid / (double)cfg->grid_lim);
        size_t nt = (size_t)g.nb * g.nb * g.nb;
        g.cap = (int)nt;  // directly addressed: one pool entry per block of the grid
        if (nt * BN * sizeof(float4) * 5 > ((size_t)64 << 30)) throw std::string("n_grid too large for the directly addressed grid (5 node arrays > 64 GB)");
        g.table = s->dalloc<int>(nt);
        g.n_slots = s->dalloc<int>(1);
        g.slot_coord = s->dalloc<int>(g.cap);
        g.flags = s->dalloc<int>(4);
        size_t pn = (size_t)g.cap * BN;
        g.acc = s->dalloc<float4>(pn);
        g.vout = s->dalloc<float4>(pn);
        g.colv = s->dalloc<float4>(pn);
        g.coln = s->dalloc<float4>(pn);
        g.mov = s->dalloc<float4>(pn);
        g.dbg_acc = nullptr;
        g.clk = s->dalloc<unsigned long long>(64 * 64);
        g.ts = nullptr;
        k_fill_int<<<1184, 256>>>(g.table, nt, -1);
        // sorted particle arrays (two copies, see resort), each with 32 records of slack: the 16-byte bulk-copy granule and the
        // lanes past the end of the last slab read in bounds
        auto alloc_recs = [&](Recs& R) {
            size_t ne = (size_t)s->Ne + 32, nt = (size_t)s->Nt + 32, nv = (size_t)s->Nv + 32;
            R.EFM = s->dalloc<int4>(ne);
            R.K0 = s->dalloc<float4>(ne); R.K1 = s->dalloc<float4>(ne); R.XE = s->dalloc<float4>(ne);
            R.C0 = s->dalloc<float4>(ne); R.C1 = s->dalloc<float4>(ne); R.SP3 = s->dalloc<float4>(ne);
            R.EV = s->dalloc<float4>(ne); R.ED1 = s->dalloc<float4>(ne); R.ED2 = s->dalloc<float4>(ne);
            R.CE = s->dalloc<int>(ne); R.CV = s->dalloc<int>(nv);
            for (int b = 0; b < 2; b++) R.D3[b] = s->dalloc<float4>(ne);
            R.TP = s->dalloc<float>(nt * KP_F); R.TS = s->dalloc<float>(nt * S_F); R.TF = s->dalloc<float>(nt * TF_F);
            R.VP = s->dalloc<float>(nv * VP_F); R.VF[0] = s->dalloc<float4>(nv); R.VF[1] = s->dalloc<float4>(nv);
            R.permE = s->dalloc<uint32_t>(s->Ne); R.permT = s->dalloc<uint32_t>(s->Nt); R.permV = s->dalloc<uint32_t>(s->Nv);
            R.invE = s->dalloc<int>(s->Ne); R.invT = s->dalloc<int>(s->Nt); R.invV = s->dalloc<int>(s->Nv);
        };
        alloc_recs(s->R);
        alloc_recs(s->R2);
        int nmax = std::max(s->Ne, std::max(s->Nt, s->Nv));
        s->ordE = s->dalloc<uint32_t>(s->Ne); s->ordT = s->dalloc<uint32_t>(s->Nt); s->ordV = s->dalloc<uint32_t>(s->Nv);
        s->o2nV = s->dalloc<int>(s->Nv);
        s->keys_in = s->dalloc<uint32_t>(nmax); s->keys_out = s->dalloc<uint32_t>(nmax); s->vals_in = s->dalloc<uint32_t>(nmax);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, s->cub_bytes, s->keys_in, s->keys_out, s->vals_in, s->R.permE, nmax, 0, 32));
        s->cub_tmp = s->dalloc<char>(s->cub_bytes);
        Canon& c = s->canon;
        int N = s->N, Nnv = s->Nnv;
        c.x = s->dalloc<float>(3 * (size_t)N); c.v = s->dalloc<float>(3 * (size_t)N); c.C = s->dalloc<float>(9 * (size_t)N);
        c.F = s->dalloc<float>(9 * (size_t)Nnv); c.Ft = s->dalloc<float>(9 * (size_t)Nnv); c.stress = s->dalloc<float>(9 * (size_t)Nnv);
        c.d = s->dalloc<float>(9 * (size_t)s->Ne); c.Rinv = s->dalloc<float>(3 * (size_t)s->Ne); c.faces = s->dalloc<float>(3 * (size_t)s->Ne);
        c.vforce = s->dalloc<float>(3 * (size_t)s->Nv);
        c.vol = s->dalloc<float>(N); c.mass = s->dalloc<float>(N); c.mu = s->dalloc<float>(N); c.lam = s->dalloc<float>(N);
        c.gamma = s->dalloc<float>(N); c.kappa = s->dalloc<float>(N); c.ys = s->dalloc<float>(N);
        {   // F = F_trial = I by default (reset_state, mpm_data_structure.py:348-360)
            std::vector<float> eye(9 * (size_t)std::max(Nnv, 1), 0.f);
            for (int i = 0; i < Nnv; i++) eye[9 * (size_t)i] = eye[9 * (size_t)i + 4] = eye[9 * (size_t)i + 8] = 1.f;
            CK(cudaMemcpy(c.F, eye.data(), 9 * (size_t)Nnv * sizeof(float), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c.Ft, eye.data(), 9 * (size_t)Nnv * sizeof(float), cudaMemcpyHostToDevice));
        }
        s->mesh_faces = s->dalloc<int>(3 * (size_t)cfg->n_mesh_f);
        s->mesh_x = s->dalloc<float>(3 * (size_t)cfg->n_mesh_v);
        s->mesh_v = s->dalloc<float>(3 * (size_t)cfg->n_mesh_v);
        s->joint_t = s->dalloc<float>(3 * (size_t)std::max(s->Nt, 1));
        s->joint_v = s->dalloc<float>(3 * (size_t)std::max(cfg->num_joint_v, 1));
        s->joint_f = s->dalloc<float>(3 * (size_t)std::max(cfg->num_joint_f, 1));
        s->d_bcs = s->dalloc<BCDesc>(MAX_BC);
        s->d_ops = s->dalloc<ParticleOp>(MAX_OPS);
        s->st = s->dalloc<StepState>(1);
        CK(cudaHostAlloc((void**)&s->h_nslots, sizeof(int), cudaHostAllocDefault));
        *s->h_nslots = 0;
        // model defaults (mpm_data_structure.py:686-715)
        s->md.material = 0; s->md.hardening = 0; s->md.friction_coeff = 0.f; s->md.alpha = 0.f;
        s->md.gx = s->md.gy = s->md.gz = 0.f; s->md.rpic = 0.f; s->md.damping = 1.1f;
        s->md.xi = 0.f; s->md.plastic_viscosity = 0.f; s->md.softening = 0.1f;
        CK(cudaFuncSetAttribute(k_p2g_elements, cudaFuncAttributeMaxDynamicSharedMemorySize, P2G_SMEM));
        CK(cudaFuncSetAttribute(k_p2g<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2G_SMEM));
        CK(cudaFuncSetAttribute(k_p2g<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + P2G_V_NW * P2G_WB));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
    } catch (const std::string& e) {
        g_create_error = e;
        for (void* p : s->allocs) cudaFree(p);
        delete s;
        return -2;
    }
    *out = s;
    return 0;
}

void mpm_destroy(MpmSolver* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    cudaDeviceSynchronize();
    destroy_graphs(s);
    destroy_sharded(s);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    if (s->h_nslots) cudaFreeHost(s->h_nslots);
    for (void* p : s->allocs) cudaFree(p);
    delete s;
}

const char* mpm_last_error(MpmSolver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int mpm_set_model(MpmSolver* s, const MpmModelParams* p) {
    API_BEGIN(s)
    // the captured substep graphs carry the model scalars as by-value kernel arguments (gravity, material, damping,
    // rpic, friction, ...): a change invalidates them, or a replay would silently step with the old parameters
    const ModelDev old = s->md;
    s->md.material = p->material; s->md.hardening = p->hardening;
    s->md.friction_coeff = p->friction_coeff; s->md.alpha = p->alpha;
    s->md.gx = p->g[0]; s->md.gy = p->g[1]; s->md.gz = p->g[2];
    s->md.rpic = p->rpic_damping; s->md.damping = p->grid_v_damping_scale;
    s->md.xi = p->xi; s->md.plastic_viscosity = p->plastic_viscosity; s->md.softening = p->softening;
    if (memcmp(&old, &s->md, sizeof(ModelDev)) != 0) {
        destroy_graphs(s);
        destroy_shard_graphs(s);
    }
    API_END(s)
}

int mpm_import_state(MpmSolver* s, const MpmParticleArrays* a, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    // bring the mirror up to date first so that a partial import keeps the other fields
    export_to_canon(s, q);
    Canon& c = s->canon;
    size_t N = s->N, Nnv = s->Nnv, Ne = s->Ne;
    auto cp = [&](float* dst, const float* src, size_t n) {
        if (src && n) CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDefault, q));
    };
    cp(c.x, a->x, 3 * N); cp(c.v, a->v, 3 * N); cp(c.C, a->C, 9 * N);
    cp(c.F, a->F, 9 * Nnv); cp(c.Ft, a->F_trial, 9 * Nnv);
    cp(c.d, a->d, 9 * Ne); cp(c.Rinv, a->R_inv, 3 * Ne); cp(c.faces, a->faces, 3 * Ne);
    cp(c.vol, a->vol, N); cp(c.mass, a->mass, N);
    cp(c.mu, a->mu, N); cp(c.lam, a->lam, N); cp(c.gamma, a->gamma, N); cp(c.kappa, a->kappa, N);
    cp(c.ys, a->yield_stress, N);
    s->have_state = true;
    s->need_sort = true;
    API_END(s)
}

int mpm_export_state(MpmSolver* s, const MpmParticleArrays* a, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (s->need_sort && s->have_state && !s->canon_stale) { /* mirror already current */ }
    export_to_canon(s, q);
    Canon& c = s->canon;
    size_t N = s->N, Nnv = s->Nnv, Ne = s->Ne, Nv = s->Nv;
    auto cp = [&](float* dst, const float* src, size_t n) {
        if (dst && n) CK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDefault, q));
    };
    cp(a->x, c.x, 3 * N); cp(a->v, c.v, 3 * N); cp(a->C, c.C, 9 * N);
    cp(a->F, c.F, 9 * Nnv); cp(a->F_trial, c.Ft, 9 * Nnv); cp(a->stress, c.stress, 9 * Nnv);
    cp(a->d, c.d, 9 * Ne); cp(a->vertex_force, c.vforce, 3 * Nv);
    cp(a->mu, c.mu, N); cp(a->lam, c.lam, N); cp(a->yield_stress, c.ys, N);
    cp(a->mass, c.mass, N); cp(a->vol, c.vol, N);
    API_END(s)
}

int mpm_set_body_mesh(MpmSolver* s, const int* faces, const float* points0, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (s->cfg.n_mesh_f == 0) throw std::string("solver was created without a body mesh");
    CK(cudaMemcpyAsync(s->mesh_faces, faces, 3 * (size_t)s->cfg.n_mesh_f * sizeof(int), cudaMemcpyDefault, q));
    if (points0) CK(cudaMemcpyAsync(s->mesh_x, points0, 3 * (size_t)s->cfg.n_mesh_v * sizeof(float), cudaMemcpyDefault, q));
    API_END(s)
}

int mpm_add_mesh_collider(MpmSolver* s, float friction) {
    API_BEGIN(s)
    if (s->cfg.n_mesh_f == 0) throw std::string("add_mesh_collider: no body mesh");
    if (s->has_collider) throw std::string("only one mesh collider is supported");
    s->has_collider = true;
    s->col_friction = friction;
    destroy_graphs(s);  // the friction coefficient is a by-value argument of the captured grid update
    destroy_shard_graphs(s);
    API_END(s)
}

int mpm_add_particle_mover(MpmSolver* s) {
    API_BEGIN(s)
    if (s->has_mover) throw std::string("only one particle mover is supported");
    s->has_mover = true;
    API_END(s)
}

static int push_bc(MpmSolver* s, const BCDesc& b) {
    if ((int)s->h_bcs.size() >= MAX_BC) { s->err = "too many grid boundary conditions"; return -2; }
    s->h_bcs.push_back(b);
    s->bcs_dirty = true;
    return 0;
}

int mpm_add_surface_collider(MpmSolver* s, const float point[3], const float normal[3], int surface_type, float friction,
                             float start_time, float end_time) {
    if (!s) return -1;
    BCDesc b{};
    b.kind = 0; b.surface_type = surface_type; b.friction = friction; b.start_time = start_time; b.end_time = end_time;
    for (int i = 0; i < 3; i++) { b.point[i] = point[i]; b.normal[i] = normal[i]; }
    return push_bc(s, b);
}
int mpm_set_velocity_on_cuboid(MpmSolver* s, const float point[3], const float size[3], const float velocity[3],
                               float start_time, float end_time, int reset) {
    if (!s) return -1;
    BCDesc b{};
    b.kind = 1; b.reset = reset; b.start_time = start_time; b.end_time = end_time;
    for (int i = 0; i < 3; i++) { b.point[i] = point[i]; b.size[i] = size[i]; b.velocity[i] = velocity[i]; }
    return push_bc(s, b);
}
int mpm_add_bounding_box(MpmSolver* s, float start_time, float end_time) {
    if (!s) return -1;
    BCDesc b{};
    b.kind = 2; b.start_time = start_time; b.end_time = end_time;
    return push_bc(s, b);
}
int mpm_enforce_grid_velocity_by_mask(MpmSolver* s, const int* mask, void* stream) {
    API_BEGIN(s)
    size_t n3 = (size_t)s->g.n * s->g.n * s->g.n;
    int* d = s->dalloc<int>(n3);
    CK(cudaMemcpyAsync(d, mask, n3 * sizeof(int), cudaMemcpyDefault, (cudaStream_t)stream));
    BCDesc b{};
    b.kind = 3; b.mask = d; b.start_time = 0.f; b.end_time = 1e30f;
    if (push_bc(s, b)) return -2;
    API_END(s)
}
}  // extern "C"
// The reference applies every impulse (pre_p2g_operations) before every velocity modifier, each group in the order
// it was added, whatever the order of the calls (mpm_solver.py:260-279): impulses go in front of the first modifier.
static void insert_particle_op(MpmSolver* s, const ParticleOp& op) {
    auto pos = s->h_ops.end();
    if (op.kind <= 1) pos = std::find_if(s->h_ops.begin(), s->h_ops.end(), [](const ParticleOp& o) { return o.kind >= 2; });
    s->h_ops.insert(pos, op);
    s->ops_dirty = true;
}
extern "C" {
int mpm_add_particle_op(MpmSolver* s, int kind, const float vec[3], const int* mask, float start_time, float end_time,
                        void* stream) {
    API_BEGIN(s)
    if ((int)s->h_ops.size() >= MAX_OPS) throw std::string("too many particle operations");
    int* d = s->dalloc<int>(s->N);
    CK(cudaMemcpyAsync(d, mask, (size_t)s->N * sizeof(int), cudaMemcpyDefault, (cudaStream_t)stream));
    ParticleOp op{};
    op.kind = kind; op.mask = d; op.start_time = start_time; op.end_time = end_time;
    for (int i = 0; i < 3; i++) op.vec[i] = vec[i];
    insert_particle_op(s, op);
    API_END(s)
}

int mpm_add_particle_rotation(MpmSolver* s, const float point[3], const float normal[3], const float axis1[3], const float axis2[3],
                              float rotation_scale, float translation_scale, const int* mask, float start_time, float end_time,
                              void* stream) {
    API_BEGIN(s)
    if ((int)s->h_ops.size() >= MAX_OPS) throw std::string("too many particle operations");
    int* d = s->dalloc<int>(s->N);
    CK(cudaMemcpyAsync(d, mask, (size_t)s->N * sizeof(int), cudaMemcpyDefault, (cudaStream_t)stream));
    ParticleOp op{};
    op.kind = 3; op.mask = d; op.start_time = start_time; op.end_time = end_time;
    for (int i = 0; i < 3; i++) { op.point[i] = point[i]; op.n[i] = normal[i]; op.h1[i] = axis1[i]; op.h2[i] = axis2[i]; }
    op.rot = rotation_scale;
    op.trans = translation_scale;
    insert_particle_op(s, op);
    API_END(s)
}

int mpm_set_time(MpmSolver* s, double t) {
    API_BEGIN(s)
    StepState h{t, 0, 0};
    CK(cudaMemcpy(s->st, &h, sizeof h, cudaMemcpyHostToDevice));
    s->host_time = t;
    API_END(s)
}

int mpm_step(MpmSolver* s, float dt, int nsub, const MpmFrameInputs* in, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (!s->have_state) throw std::string("mpm_step before mpm_import_state");
    MpmFrameInputs none{};
    if (!in) in = &none;
    if (in->n_joint_t > s->Nt) throw std::string("n_joint_t exceeds the number of traditional particles");
    upload_lists(s, q);
    SubstepArgs a{};
    a.dt = dt;
    a.collider = s->has_collider;
    a.advance_mesh = in->mesh_x != nullptr && nsub > 1;
    // the mover runs only when BOTH joint_verts_v and joint_faces_v are given (mpm_solver.py:421)
    a.mover = s->has_mover && in->joint_verts_v && in->joint_faces_v;
    a.njt = (a.mover && in->joint_traditional_v && in->n_joint_t > 0) ? in->n_joint_t : 0;
    const int mv3 = 3 * s->cfg.n_mesh_v;
    StageInputs si{};
    int nsi = 0, most = 0;
    auto stage = [&](float* dst, const float* src, int n) {
        if (!src || n <= 0) return;
        if (in->device_inputs) { si.src[nsi] = src; si.dst[nsi] = dst; si.n[nsi] = n; nsi++; most = std::max(most, n); }
        else CK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDefault, q));
    };
    stage(s->mesh_x, in->mesh_x, mv3);
    stage(s->mesh_v, in->mesh_v, mv3);
    if (a.mover) {
        stage(s->joint_v, in->joint_verts_v, 3 * s->cfg.num_joint_v);
        stage(s->joint_f, in->joint_faces_v, 3 * s->cfg.num_joint_f);
        stage(s->joint_t, in->joint_traditional_v, 3 * a.njt);
    }
    if (in->device_inputs) k_stage_inputs<<<std::max(1, std::min(cdiv(most, 256), 148)), 256, 0, q>>>(si, s->st);  // + substep counter reset
    else k_reset_k<<<1, 1, 0, q>>>(s->st);
    s->launches++;
    int left = nsub;
    while (left > 0) {
        if (s->need_sort || s->since_sort >= s->resort_interval) resort(s, q);
        int chunk = std::min(left, s->resort_interval - s->since_sort);
        run_substeps(s, a, chunk, q, nsub == 1);
        s->since_sort += chunk;
        s->n_substeps += chunk;
        s->canon_stale = true;
        for (int k = 0; k < chunk; k++) s->host_time += (double)dt;
        left -= chunk;
    }
    CK(cudaGetLastError());
    API_END(s)
}

// ---- sharded stepping (one substep, split around the exchange of the shared grid blocks)
static void begin_half_step(MpmSolver* s, const MpmFrameInputs* in, SubstepArgs& a, float dt, cudaStream_t q) {
    if (!s->have_state) throw std::string("mpm_step_* before mpm_import_state");
    a.dt = dt;
    a.collider = s->has_collider;
    a.advance_mesh = false;
    a.mover = s->has_mover && in && in->joint_verts_v && in->joint_faces_v;
    a.njt = 0;
}

int mpm_step_scatter(MpmSolver* s, float dt, const MpmFrameInputs* in, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    MpmFrameInputs none{};
    if (!in) in = &none;
    SubstepArgs a{};
    begin_half_step(s, in, a, dt, q);
    if (in->n_joint_t > s->Nt) throw std::string("n_joint_t exceeds the number of traditional particles");
    upload_lists(s, q);
    size_t mv3 = 3 * (size_t)s->cfg.n_mesh_v * sizeof(float);
    if (in->mesh_x && mv3) CK(cudaMemcpyAsync(s->mesh_x, in->mesh_x, mv3, cudaMemcpyDefault, q));
    if (in->mesh_v && mv3) CK(cudaMemcpyAsync(s->mesh_v, in->mesh_v, mv3, cudaMemcpyDefault, q));
    if (a.mover) {
        if (s->cfg.num_joint_v) CK(cudaMemcpyAsync(s->joint_v, in->joint_verts_v, 3 * (size_t)s->cfg.num_joint_v * sizeof(float), cudaMemcpyDefault, q));
        if (s->cfg.num_joint_f) CK(cudaMemcpyAsync(s->joint_f, in->joint_faces_v, 3 * (size_t)s->cfg.num_joint_f * sizeof(float), cudaMemcpyDefault, q));
        if (in->joint_traditional_v && in->n_joint_t > 0) {  // this rank's share of the pinned tail (sharding.local_joint_traditional)
            a.njt = in->n_joint_t;
            CK(cudaMemcpyAsync(s->joint_t, in->joint_traditional_v, 3 * (size_t)a.njt * sizeof(float), cudaMemcpyDefault, q));
        }
    }
    if (s->need_sort || s->since_sort >= s->resort_interval) resort(s, q);
    s->pending_mover = a.mover;
    s->pending_njt = a.njt;
    launch_substep(s, a, q, HALF_SCATTER);
    CK(cudaGetLastError());
    API_END(s)
}

int mpm_step_gather(MpmSolver* s, float dt, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    SubstepArgs a{};
    begin_half_step(s, nullptr, a, dt, q);
    a.mover = s->pending_mover;
    a.njt = s->pending_njt;
    launch_substep(s, a, q, HALF_GATHER);
    s->since_sort++;
    s->n_substeps++;
    s->canon_stale = true;
    s->host_time += (double)dt;
    CK(cudaGetLastError());
    API_END(s)
}

// nsub sharded substeps driven from C: per substep the host only enqueues ~9 kernels and calls `exchange`
// (the caller's all-reduce on its communicator, e.g. torch.distributed / NCCL) once
int mpm_step_sharded(MpmSolver* s, float dt, int nsub, const MpmFrameInputs* in, float* buf, int refresh,
                     MpmExchangeFn exchange, MpmRebuildFn rebuild, void* ctx, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    MpmFrameInputs none{};
    if (!in) in = &none;
    SubstepArgs a{};
    begin_half_step(s, in, a, dt, q);
    if (in->n_joint_t > s->Nt) throw std::string("n_joint_t exceeds the number of traditional particles");
    upload_lists(s, q);
    size_t mv3 = 3 * (size_t)s->cfg.n_mesh_v * sizeof(float);
    if (in->mesh_x && mv3) CK(cudaMemcpyAsync(s->mesh_x, in->mesh_x, mv3, cudaMemcpyDefault, q));
    if (in->mesh_v && mv3) CK(cudaMemcpyAsync(s->mesh_v, in->mesh_v, mv3, cudaMemcpyDefault, q));
    if (a.mover) {
        if (s->cfg.num_joint_v) CK(cudaMemcpyAsync(s->joint_v, in->joint_verts_v, 3 * (size_t)s->cfg.num_joint_v * sizeof(float), cudaMemcpyDefault, q));
        if (s->cfg.num_joint_f) CK(cudaMemcpyAsync(s->joint_f, in->joint_faces_v, 3 * (size_t)s->cfg.num_joint_f * sizeof(float), cudaMemcpyDefault, q));
        if (in->joint_traditional_v && in->n_joint_t > 0) {  // this rank's share of the pinned tail (sharding.local_joint_traditional)
            a.njt = in->n_joint_t;
            CK(cudaMemcpyAsync(s->joint_t, in->joint_traditional_v, 3 * (size_t)a.njt * sizeof(float), cudaMemcpyDefault, q));
        }
    }
    a.advance_mesh = in->mesh_x != nullptr && nsub > 1;  // substep k sees mesh_x + dt*k*mesh_v (device-side counter)
    k_reset_k<<<1, 1, 0, q>>>(s->st);
    s->launches++;
    for (int k = 0; k < nsub; k++) {
        if (s->need_sort || s->since_sort >= s->resort_interval) resort(s, q);
        launch_substep(s, a, q, HALF_SCATTER);
        if (s->shared_age < 0 || s->shared_age >= refresh) {
            if (rebuild(ctx) != 0) throw std::string("shared-block rebuild callback failed");
            s->shared_age = 0;
        }
        if (s->n_shared) {
            k_shared_pack<<<cdiv((long long)s->n_shared * BN, 256), 256, 0, q>>>(s->g, s->d_shared, s->n_shared, nullptr, (float4*)buf);
            if (exchange(ctx, buf, s->n_shared * BN * 8) != 0) throw std::string("exchange callback failed");
            k_shared_unpack<<<cdiv((long long)s->n_shared * BN, 256), 256, 0, q>>>(s->g, s->d_shared, s->n_shared, nullptr, (const float4*)buf);
            s->launches += 2;
        }
        launch_substep(s, a, q, HALF_GATHER);
        s->shared_age++;
        s->since_sort++;
        s->n_substeps++;
        s->host_time += (double)dt;
    }
    s->canon_stale = true;
    CK(cudaGetLastError());
    API_END(s)
}

}  // extern "C" (re-opened after the sharded helpers)

// ---- in-graph sharded stepping over this solver's own NCCL communicator
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib) return &api;
    // the soname torch's own NCCL carries: inside a torch process this resolves to the library already loaded
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    auto sym = [&](const char* name) {
        void* p = dlsym(lib, name);
        if (!p) throw std::string("libnccl.so.2 lacks ") + name;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.lib = lib;
    return &api;
}
#define NCK(call)                                                                            \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess) throw std::string(#call " failed: ") + nccl_api()->GetErrorString(r_); \
    } while (0)

struct ShardGraphKey {
    float dt;
    int collider, mover, advance_mesh, cur, n_bc, n_ops, xcap, len, rbuf, njt;
    const void* shared_ptr;
    bool operator==(const ShardGraphKey& o) const { return memcmp(this, &o, sizeof(ShardGraphKey)) == 0; }
};
struct ShardGraph {
    ShardGraphKey key;
    cudaGraphExec_t exec;
    int launches;
};
std::vector<ShardGraph>& shard_graphs(MpmSolver* s) {
    if (!s->shard_graphs) s->shard_graphs = new std::vector<ShardGraph>();
    return *static_cast<std::vector<ShardGraph>*>(s->shard_graphs);
}
// ---- the few collectives the sharded path needs, over either transport: the solver's own NCCL communicator
// (mpm_attach_comm), or a blocking HOST all-gather supplied by the caller (mpm_attach_host_comm: gloo / MPI ranks, also
// ranks that share one GPU, which NCCL refuses).  The host transport synchronises the stream.
bool have_comm(const MpmSolver* s) { return s->comm != nullptr || s->host_ag != nullptr; }
void host_allgather(MpmSolver* s, const void* send, void* recv, size_t nbytes) {
    if (s->host_ag(s->host_ctx, send, recv, (int)nbytes) != 0) throw std::string("host all-gather callback failed");
}
// every rank's `nbytes` host bytes, in rank order
void allgather_host_bytes(MpmSolver* s, const void* send, void* recv, size_t nbytes, cudaStream_t q) {
    if (s->host_ag) { host_allgather(s, send, recv, nbytes); return; }
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) =
        (decltype(AllGather))dlsym(nccl_api()->lib, "ncclAllGather");
    if (!AllGather) throw std::string("libnccl.so.2 lacks ncclAllGather");
    const size_t need = (size_t)(s->comm_size + 1) * nbytes;
    if (need > s->ag_bytes) {  // small, rarely used (set-up traffic): grown, never shrunk
        s->ag_dev = s->dalloc<unsigned char>(need);
        s->ag_bytes = need;
    }
    unsigned char* d = s->ag_dev;
    CK(cudaMemcpyAsync(d + (size_t)s->comm_size * nbytes, send, nbytes, cudaMemcpyHostToDevice, q));
    NCK(AllGather(d + (size_t)s->comm_size * nbytes, d, nbytes, ncclUint8, s->comm, q));
    CK(cudaMemcpyAsync(recv, d, (size_t)s->comm_size * nbytes, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
}
// in-place sum over the ranks of a DEVICE buffer: bytes (the block marks) or floats (the packed shared blocks; the host
// transport adds the ranks' parts in rank order, so every rank gets bit-identical sums)
template <typename T>
void allreduce_sum_device(MpmSolver* s, T* d, size_t n, cudaStream_t q) {
    if (s->comm) {
        NCK(nccl_api()->AllReduce(d, d, n, sizeof(T) == 1 ? ncclUint8 : ncclFloat, ncclSum, s->comm, q));
        return;
    }
    const size_t bytes = n * sizeof(T);
    s->host_stage.resize((size_t)(s->comm_size + 1) * bytes);
    unsigned char* mine = s->host_stage.data() + (size_t)s->comm_size * bytes;
    CK(cudaMemcpyAsync(mine, d, bytes, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    host_allgather(s, mine, s->host_stage.data(), bytes);
    T* acc = reinterpret_cast<T*>(mine);
    const T* all = reinterpret_cast<const T*>(s->host_stage.data());
    for (size_t i = 0; i < n; i++) {
        T v = all[i];
        for (int r = 1; r < s->comm_size; r++) v = (T)(v + all[(size_t)r * n + i]);
        acc[i] = v;
    }
    CK(cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, q));
    CK(cudaStreamSynchronize(q));
}

// one sharded substep on stream q.  Peer-to-peer: the plain chain with the exchange fused into the grid update;
// otherwise scatter half, pack, all-reduce, unpack, gather half
void sharded_substep(MpmSolver* s, const SubstepArgs& a, cudaStream_t q) {
    if (s->p2p_ready) {
        PeerArea P = s->peer;
        P.mapA = s->d_mapA; P.mapM = s->d_mapM; P.memA = s->d_memA; P.memM = s->d_memM;
        P.listA = s->d_shared; P.listM = s->d_sharedM; P.nA = s->d_n_shared; P.nM = s->d_nM;
        P.capA = s->xcap_blocks; P.capM = s->xcapM;
        s->gu_peer = P;
        launch_substep(s, a, q, HALF_SCATTER | HALF_GATHER);
        s->gu_peer = PeerArea{};
        return;
    }
    launch_substep(s, a, q, HALF_SCATTER);
    const int ctas = std::max(1, std::min(cdiv((long long)(s->xcap_blocks + s->xcapM) * BN, 256), 148 * 4));
    const SharedLists L{s->d_shared, s->d_n_shared, s->d_sharedM, s->d_nM, s->xcap_blocks, s->xcapM};
    k_shared_pack2<<<ctas, 256, 0, q>>>(s->g, L, (float4*)s->xbuf);
    allreduce_sum_device(s, s->xbuf, (size_t)(s->xcap_blocks + s->xcapM) * BN * 4, q);
    k_shared_unpack2<<<ctas, 256, 0, q>>>(s->g, L, (const float4*)s->xbuf);
    s->launches += 3;
    launch_substep(s, a, q, HALF_GATHER);
}
}  // namespace

// Rebuild of the shared-block list entirely on the stream: mark, byte-sum over the ranks, ordered compaction.
// The host only learns the count one rebuild later (pinned mirror), which is enough to grow the buffers in time.
static void mark_potential(MpmSolver* s, int margin, cudaStream_t q) {
    const size_t nt = (size_t)s->g.nb * s->g.nb * s->g.nb;
    if (!s->d_mark) {
        s->d_mark = s->dalloc<unsigned char>(2 * nt);  // [rank count | joint-rank count]
        s->h_mark.resize(nt);
    }
    unsigned char* mj = s->d_mark + nt;
    CK(cudaMemsetAsync(s->d_mark, 0, 2 * nt, q));
    const int bit = (have_comm(s) && s->comm_size <= 8) ? (1 << s->comm_rank) : 1;  // rank set for <= 8 ranks, else a count
    if (s->Ne) k_mark_potential<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->g, s->Ne, (const float*)s->R.XE, 4, margin, s->d_mark, mj, s->R.permE, s->cfg.num_joint_f, bit);
    // any traditional particle may belong to the pinned tail (its length changes from call to call, run_demo.py:524)
    if (s->Nt) k_mark_potential<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->g, s->Nt, s->R.TP, KP_F, margin, s->d_mark, mj, s->R.permT, s->Nt, bit);
    if (s->Nv) k_mark_potential<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->g, s->Nv, s->R.VP, VP_F, margin, s->d_mark, mj, s->R.permV, s->cfg.num_joint_v, bit);
    s->launches += 3;
}
// (Re)allocate this rank's receive area for the current capacities and map every peer's (CUDA IPC; the handles travel
// through ncclAllGather).  Collective: every rank resizes at the same rebuild because the lists are global.  Any
// failure (no peer access, more than 8 ranks, MPM_B200_P2P=0) leaves the ncclAllReduce exchange in place.
static void close_peer_areas(MpmSolver* s) {
    for (int r = 0; r < 8; r++)
        if (s->peer_open[r]) { cudaIpcCloseMemHandle(s->peer_open[r]); s->peer_open[r] = nullptr; }
    if (s->peer_local) { cudaFree(s->peer_local); s->peer_local = nullptr; }
    s->p2p_ready = false;
}
static void setup_peer_areas(MpmSolver* s, cudaStream_t q) {
    const int n = s->comm_size;
    if (!s->use_p2p || !have_comm(s) || n < 2 || n > 8) return;
    CK(cudaStreamSynchronize(q));
    // every rank has finished every exchange that used the old areas (it is inside this collective).  Imported
    // mappings are closed first; the old local area is freed only after the all-gather below, i.e. after every peer
    // has closed its mapping of it
    for (int r = 0; r < 8; r++)
        if (s->peer_open[r]) { cudaIpcCloseMemHandle(s->peer_open[r]); s->peer_open[r] = nullptr; }
    unsigned char* old_local = s->peer_local;
    s->peer_local = nullptr;
    s->p2p_ready = false;
    const size_t slot = (size_t)(s->xcap_blocks + s->xcapM) * BN * 32;  // flagged parts: 32 bytes per node
    const size_t total = 2 * (size_t)n * slot;                            // [parity][sender]
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
    Msg mine{};
    mine.ok = cudaMalloc(&s->peer_local, total) == cudaSuccess && cudaMemset(s->peer_local, 0, total) == cudaSuccess &&
              cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&mine.h, s->peer_local) == cudaSuccess;
    cudaGetLastError();
    // the handles travel over whichever transport the solver has (NCCL all-gather or the caller's host all-gather)
    std::vector<Msg> all(n);
    allgather_host_bytes(s, &mine, all.data(), sizeof(Msg), q);
    if (old_local) cudaFree(old_local);
    bool ok = true;
    for (int r = 0; r < n; r++) ok = ok && all[r].ok;
    PeerArea P{};
    for (int r = 0; r < n && ok; r++) {
        if (r == s->comm_rank) { P.base[r] = s->peer_local; continue; }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        s->peer_open[r] = ptr;
        P.base[r] = (unsigned char*)ptr;
    }
    // agree on the outcome: one rank failing to map a peer must switch every rank to the all-reduce exchange
    int h_ok = ok ? 1 : 0;
    std::vector<int> oks(n);
    allgather_host_bytes(s, &h_ok, oks.data(), sizeof(int), q);
    for (int r = 0; r < n; r++) h_ok = std::min(h_ok, oks[r]);
    if (!h_ok) { close_peer_areas(s); return; }
    P.slot_bytes = slot;
    P.rank = s->comm_rank;
    P.nranks = n;
    unsigned* ctr = s->dalloc<unsigned>(4);  // zeroed: epoch, push counter, pull counter
    P.epoch = ctr;
    P.counter = ctr + 1;
    s->peer = P;
    s->p2p_ready = true;
}

static void rebuild_shared_on_device(MpmSolver* s, int margin, cudaStream_t q) {
    const size_t nt = (size_t)s->g.nb * s->g.nb * s->g.nb;
    cub::CountingInputIterator<int> it(0);
    if (!s->d_sel) {
        s->d_sel = s->dalloc<int>(2 * nt);
        s->d_n_shared = s->d_n_shared ? s->d_n_shared : s->dalloc<int>(1);
        s->d_nM = s->dalloc<int>(1);
        CK(cudaMallocHost(&s->h_nshared, 2 * sizeof(int)));
        s->h_nshared[0] = s->h_nshared[1] = -1;
        CK(cub::DeviceSelect::If(nullptr, s->sel_bytes, it, s->d_sel, s->d_n_shared, (int)nt, SharedPred{s->d_mark, s->d_mark, 1}, q));
        s->sel_tmp = s->dalloc<unsigned char>(s->sel_bytes);
    }
    const bool first = s->h_nshared[0] < 0;
    // the lists (as of the previous rebuild) near their capacities: re-size from the actual counts below
    bool resize = first || s->h_nshared[0] * 10 > s->xcap_blocks * 9 || s->h_nshared[1] * 10 > s->xcapM * 9;
    mark_potential(s, margin, q);
    allreduce_sum_device(s, s->d_mark, 2 * nt, q);  // a sum of distinct rank bits is the member set
    const int bits = s->comm_size <= 8 ? 1 : 0;
    size_t tmp = s->sel_bytes;
    CK(cub::DeviceSelect::If(s->sel_tmp, tmp, it, s->d_sel, s->d_n_shared, (int)nt, SharedPred{s->d_mark, nullptr, bits}, q));
    tmp = s->sel_bytes;
    CK(cub::DeviceSelect::If(s->sel_tmp, tmp, it, s->d_sel + nt, s->d_nM, (int)nt, SharedPred{s->d_mark, s->d_mark + nt, bits}, q));
    s->launches += 3;
    if (resize) {  // the only synchronising rebuilds
        int n[2] = {0, 0};
        CK(cudaMemcpyAsync(&n[0], s->d_n_shared, sizeof(int), cudaMemcpyDeviceToHost, q));
        CK(cudaMemcpyAsync(&n[1], s->d_nM, sizeof(int), cudaMemcpyDeviceToHost, q));
        CK(cudaStreamSynchronize(q));
        // 100 % head-room: a re-size synchronises, re-maps the receive areas and re-captures the windows (2-8 ms), and the
        // lists of a garment that is still settling grow by a quarter within a few hundred substeps
        auto room = [](int k) { return (2 * k + 32 + 31) / 32 * 32; };
        if (getenv("MPM_B200_TRACE")) fprintf(stderr, "[mpm_b200 rank %d] substep %lld: shared lists resized (%d, %d blocks; capacities were %d, %d)\n", s->comm_rank, (long long)s->n_substeps, n[0], n[1], s->xcap_blocks, s->xcapM);
        s->xcap_blocks = room(n[0]);
        s->xcapM = room(n[1]);
        s->xbuf = s->dalloc<float>((size_t)(s->xcap_blocks + s->xcapM) * BN * 4);
        s->shared_cap = s->xcap_blocks;
        s->d_shared = s->dalloc<int>(s->xcap_blocks);
        s->d_sharedM = s->dalloc<int>(s->xcapM);
        s->d_memA = s->dalloc<unsigned char>(s->xcap_blocks);
        s->d_memM = s->dalloc<unsigned char>(s->xcapM);
        setup_peer_areas(s, q);
    }
    if (!s->d_mapA) { s->d_mapA = s->dalloc<int>(nt); s->d_mapM = s->dalloc<int>(nt); }
    k_fill_int<<<std::min(cdiv((long long)nt, 256), 1184), 256, 0, q>>>(s->d_mapA, nt, -1);
    k_fill_int<<<std::min(cdiv((long long)nt, 256), 1184), 256, 0, q>>>(s->d_mapM, nt, -1);
    k_shared_coords<<<std::max(1, std::min(cdiv(s->xcap_blocks, 256), 64)), 256, 0, q>>>(s->g, s->d_sel, s->d_n_shared, s->xcap_blocks, s->d_shared, s->d_mark, s->d_memA, s->d_mapA);
    k_shared_coords<<<std::max(1, std::min(cdiv(s->xcapM, 256), 64)), 256, 0, q>>>(s->g, s->d_sel + nt, s->d_nM, s->xcapM, s->d_sharedM, s->d_mark, s->d_memM, s->d_mapM);
    CK(cudaMemcpyAsync(&s->h_nshared[0], s->d_n_shared, sizeof(int), cudaMemcpyDeviceToHost, q));
    CK(cudaMemcpyAsync(&s->h_nshared[1], s->d_nM, sizeof(int), cudaMemcpyDeviceToHost, q));
    s->launches += 2;
    s->n_rebuilds++;
}

static void destroy_shard_graphs(MpmSolver* s) {
    if (s->shard_graphs) {
        auto& v = shard_graphs(s);
        for (auto& e : v) cudaGraphExecDestroy(e.exec);
        delete &v;
        s->shard_graphs = nullptr;
    }
}
static void destroy_sharded(MpmSolver* s) {
    if (s->h_nshared) { cudaFreeHost(s->h_nshared); s->h_nshared = nullptr; }
    close_peer_areas(s);
    destroy_shard_graphs(s);
    if (s->comm) {
        nccl_api()->CommDestroy(s->comm);
        s->comm = nullptr;
    }
}

extern "C" {

int mpm_comm_unique_id(char* out128) {
    try {
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        NCK(nccl_api()->GetUniqueId(&id));
        memcpy(out128, &id, sizeof id);
    } catch (const std::string& e) {
        g_create_error = e;
        return -2;
    }
    return 0;
}

int mpm_attach_comm(MpmSolver* s, const char* id128, int rank, int nranks) {
    API_BEGIN(s)
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    if (s->comm) { nccl_api()->CommDestroy(s->comm); s->comm = nullptr; }
    NCK(nccl_api()->CommInitRank(&s->comm, nranks, id, rank));
    s->comm_rank = rank;
    s->comm_size = nranks;
    // connection set-up happens on first use: do it here, outside any stream capture
    float* tmp = s->dalloc<float>(256);
    NCK(nccl_api()->AllReduce(tmp, tmp, 256, ncclFloat, ncclSum, s->comm, 0));
    CK(cudaStreamSynchronize(0));
    API_END(s)
}

int mpm_attach_host_comm(MpmSolver* s, int rank, int nranks, MpmHostAllGatherFn allgather, void* ctx) {
    API_BEGIN(s)
    if (!allgather || nranks < 1 || rank < 0 || rank >= nranks) throw std::string("mpm_attach_host_comm: bad arguments");
    if (s->comm) { nccl_api()->CommDestroy(s->comm); s->comm = nullptr; }
    s->host_ag = allgather;
    s->host_ctx = ctx;
    s->comm_rank = rank;
    s->comm_size = nranks;
    API_END(s)
}

int mpm_step_sharded_nccl(MpmSolver* s, float dt, int nsub, const MpmFrameInputs* in, int refresh, int margin, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (!have_comm(s)) throw std::string("mpm_step_sharded_nccl before mpm_attach_comm / mpm_attach_host_comm");
    MpmFrameInputs none{};
    if (!in) in = &none;
    SubstepArgs a{};
    begin_half_step(s, in, a, dt, q);
    if (in->n_joint_t > s->Nt) throw std::string("n_joint_t exceeds the number of traditional particles");
    upload_lists(s, q);
    size_t mv3 = 3 * (size_t)s->cfg.n_mesh_v * sizeof(float);
    if (in->mesh_x && mv3) CK(cudaMemcpyAsync(s->mesh_x, in->mesh_x, mv3, cudaMemcpyDefault, q));
    if (in->mesh_v && mv3) CK(cudaMemcpyAsync(s->mesh_v, in->mesh_v, mv3, cudaMemcpyDefault, q));
    if (a.mover) {
        if (s->cfg.num_joint_v) CK(cudaMemcpyAsync(s->joint_v, in->joint_verts_v, 3 * (size_t)s->cfg.num_joint_v * sizeof(float), cudaMemcpyDefault, q));
        if (s->cfg.num_joint_f) CK(cudaMemcpyAsync(s->joint_f, in->joint_faces_v, 3 * (size_t)s->cfg.num_joint_f * sizeof(float), cudaMemcpyDefault, q));
        if (in->joint_traditional_v && in->n_joint_t > 0) {  // this rank's share of the pinned tail (sharding.local_joint_traditional)
            a.njt = in->n_joint_t;
            CK(cudaMemcpyAsync(s->joint_t, in->joint_traditional_v, 3 * (size_t)a.njt * sizeof(float), cudaMemcpyDefault, q));
        }
    }
    a.advance_mesh = in->mesh_x != nullptr && nsub > 1;
    k_reset_k<<<1, 1, 0, q>>>(s->st);
    s->launches++;
    constexpr int W = 16;  // substeps per captured window (even: the direction ping-pong returns to its start)
    int left = nsub;
    while (left > 0) {
        if (s->need_sort || s->since_sort >= s->resort_interval) resort(s, q);
        if (s->shared_age < 0 || s->shared_age >= refresh) {
            rebuild_shared_on_device(s, margin, q);
            s->shared_age = 0;
        }
        const int room = std::min({left, refresh - s->shared_age, s->resort_interval - s->since_sort});
        int done = 1;
        if (s->use_graphs && room >= W && (s->comm || s->p2p_ready)) {  // a host-transport all-reduce cannot be captured
            ShardGraphKey key{};
            key.dt = a.dt; key.collider = a.collider; key.mover = a.mover; key.advance_mesh = a.advance_mesh; key.cur = s->cur;
            key.n_bc = (int)s->h_bcs.size(); key.n_ops = (int)s->h_ops.size(); key.xcap = s->xcap_blocks * 65536 + s->xcapM; key.len = W + (s->p2p_ready ? 1000 : 0);
            key.shared_ptr = s->d_shared; key.rbuf = s->rbuf; key.njt = a.njt;
            auto& cache = shard_graphs(s);
            ShardGraph* hit = nullptr;
            for (auto& e : cache) if (e.key == key) hit = &e;
            if (!hit) {
                if (!s->cap_stream) CK(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
                const int before = s->launches;
                cudaGraph_t graph;
                if (getenv("MPM_B200_TRACE")) fprintf(stderr, "[mpm_b200 rank %d] substep %lld: capturing a sharded window (graph %zu)\n", s->comm_rank, (long long)s->n_substeps, cache.size());
                CK(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
                for (int i = 0; i < W; i++) sharded_substep(s, a, s->cap_stream);
                CK(cudaStreamEndCapture(s->cap_stream, &graph));
                ShardGraph e;
                e.key = key;
                e.launches = s->launches - before;
                s->launches = before;
                CK(cudaGraphInstantiate(&e.exec, graph, 0));
                cudaGraphDestroy(graph);
                cache.push_back(e);
                hit = &cache.back();
            }
            CK(cudaGraphLaunch(hit->exec, q));
            s->launches += hit->launches;
            s->have_prev = true;
            done = W;
        } else {
            sharded_substep(s, a, q);
        }
        s->shared_age += done;
        s->since_sort += done;
        s->n_substeps += done;
        for (int k = 0; k < done; k++) s->host_time += (double)dt;
        left -= done;
    }
    s->canon_stale = true;
    CK(cudaGetLastError());
    API_END(s)
}

// Timeline probe of the SHARDED chain (collective: every rank calls it with the same n): n <= 32 sharded substeps launched
// eagerly with the stamps on; out[n][10][2] as mpm_measure_timeline plus ids 8 (the push phase at the head of
// k_grid_update<true>) and 9 (its pass over the shared nodes, incl. the wait for the members' parts).  Requires the peer-to-peer exchange.
int mpm_measure_timeline_sharded(MpmSolver* s, float dt, int n, const MpmFrameInputs* in, long long* out, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (n < 1 || n > 32) throw std::string("mpm_measure_timeline_sharded: 1 <= n <= 32");
    if (!s->p2p_ready) throw std::string("mpm_measure_timeline_sharded needs the peer-to-peer exchange (step once first)");
    MpmFrameInputs none{};
    if (!in) in = &none;
    SubstepArgs a{};
    begin_half_step(s, in, a, dt, q);
    const size_t per = 2 * TS_KERNELS_SHARDED;
    std::vector<unsigned long long> h((size_t)n * per);
    for (size_t i = 0; i < h.size(); i++) h[i] = (i & 1) ? 0ull : ~0ull;
    unsigned long long* d = s->dalloc<unsigned long long>(h.size());
    CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, q));
    // no re-sort, no shared-list rebuild inside the probe: n stays far below both intervals' slack (the caller steps
    // normally before and after)
    for (int i = 0; i < n; i++) {
        s->g.ts = d + (size_t)i * per;
        sharded_substep(s, a, q);
    }
    s->g.ts = nullptr;
    s->shared_age += n;
    s->since_sort += n;
    s->n_substeps += n;
    s->canon_stale = true;
    for (int k = 0; k < n; k++) s->host_time += (double)dt;
    CK(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < h.size(); i += 2) t0 = std::min(t0, h[i]);
    for (size_t i = 0; i < h.size(); i += 2) {
        const bool ran = h[i] != ~0ull;
        out[i] = ran ? (long long)(h[i] - t0) : -1;
        out[i + 1] = ran ? (long long)(h[i + 1] - t0) : -1;
    }
    API_END(s)
}

int mpm_shared_mode(MpmSolver* s) { return !s ? -1 : (s->p2p_ready ? 2 : (s->comm ? 1 : (s->host_ag ? 3 : 0))); }

int mpm_shared_info(MpmSolver* s, int* n_shared, int* cap_blocks, int* n_rebuilds, void* stream) {
    API_BEGIN(s)
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (n_shared) *n_shared = s->h_nshared ? s->h_nshared[0] : s->n_shared;
    if (cap_blocks) *cap_blocks = (s->xcap_blocks + s->xcapM + 1) / 2;  // in units of 512 floats (the callback path's block payload)
    if (n_rebuilds) *n_rebuilds = s->n_rebuilds;
    API_END(s)
}

}  // extern "C"

extern "C" {
int mpm_get_active_blocks(MpmSolver* s, int* coords, int cap, int* n, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (s->need_sort && s->have_state) resort(s, q);
    int ns = 0;
    CK(cudaMemcpyAsync(&ns, s->g.n_slots, sizeof(int), cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    ns = std::min(ns, s->g.cap);
    *n = ns;
    if (coords && cap > 0) CK(cudaMemcpy(coords, s->g.slot_coord, (size_t)std::min(ns, cap) * sizeof(int), cudaMemcpyDefault));
    API_END(s)
}

int mpm_get_potential_blocks(MpmSolver* s, int margin, int* coords, int cap, int* n, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (s->need_sort && s->have_state) resort(s, q);
    const size_t nt = (size_t)s->g.nb * s->g.nb * s->g.nb;
    mark_potential(s, margin, q);
    CK(cudaMemcpyAsync(s->h_mark.data(), s->d_mark, nt, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    int cnt = 0;
    const int nb = s->g.nb;
    for (size_t t = 0; t < nt; t++)
        if (s->h_mark[t]) {
            if (coords && cnt < cap) {
                int bz = (int)(t % nb), by = (int)((t / nb) % nb), bx = (int)(t / ((size_t)nb * nb));
                coords[cnt] = bx | (by << 10) | (bz << 20);
            }
            cnt++;
        }
    *n = cnt;
    API_END(s)
}

int mpm_set_shared_blocks(MpmSolver* s, const int* coords, int n, void* stream) {
    API_BEGIN(s)
    if (n > s->shared_cap) {
        s->shared_cap = std::max(2 * n, 1024);
        s->d_shared = s->dalloc<int>(s->shared_cap);
    }
    if (n) CK(cudaMemcpyAsync(s->d_shared, coords, (size_t)n * sizeof(int), cudaMemcpyDefault, (cudaStream_t)stream));
    s->n_shared = n;
    if (!s->d_n_shared) s->d_n_shared = s->dalloc<int>(1);
    CK(cudaMemcpyAsync(s->d_n_shared, &s->n_shared, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    API_END(s)
}

int mpm_shared_pack(MpmSolver* s, float* buf, void* stream) {
    API_BEGIN(s)
    if (s->n_shared) {
        k_shared_pack<<<cdiv((long long)s->n_shared * BN, 256), 256, 0, (cudaStream_t)stream>>>(s->g, s->d_shared, s->n_shared, nullptr, (float4*)buf);
        s->launches++;
    }
    API_END(s)
}

int mpm_shared_unpack(MpmSolver* s, const float* buf, void* stream) {
    API_BEGIN(s)
    if (s->n_shared) {
        k_shared_unpack<<<cdiv((long long)s->n_shared * BN, 256), 256, 0, (cudaStream_t)stream>>>(s->g, s->d_shared, s->n_shared, nullptr, (const float4*)buf);
        s->launches++;
    }
    API_END(s)
}

// n substeps (<= 64) launched eagerly (same kernels, same programmatic-dependent-launch chain as mpm_step) with the
// timeline probe on; out[n][TS_KERNELS][2] = first-CTA start / last-CTA end in ns relative to the first stamp, -1 for
// kernels that did not run.  Synchronises.
int mpm_measure_timeline(MpmSolver* s, float dt, int n, const MpmFrameInputs* in, long long* out, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (n < 1 || n > 64) throw std::string("mpm_measure_timeline: 1 <= n <= 64");
    if (!s->have_state) throw std::string("mpm_measure_timeline before mpm_import_state");
    MpmFrameInputs none{};
    if (!in) in = &none;
    SubstepArgs a{};
    begin_half_step(s, in, a, dt, q);
    a.njt = 0;
    a.advance_mesh = false;  // the body stays where the last step left it
    const size_t per = 2 * TS_KERNELS;
    std::vector<unsigned long long> h((size_t)n * per);
    for (size_t i = 0; i < h.size(); i++) h[i] = (i & 1) ? 0ull : ~0ull;
    unsigned long long* d = s->dalloc<unsigned long long>(h.size());
    CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, q));
    if (s->need_sort) resort(s, q);
    const bool prof = s->profiling;
    s->profiling = false;
    if (s->use_graphs && n % 2 == 0) {  // as mpm_step runs them: one captured graph, replayed (twice: the second is timed)
        if (!s->cap_stream) CK(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
        const int before = s->launches;
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        CK(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < n; i++) {
            s->g.ts = d + (size_t)i * per;
            launch_substep(s, a, s->cap_stream);
        }
        CK(cudaStreamEndCapture(s->cap_stream, &graph));
        s->g.ts = nullptr;
        CK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        CK(cudaGraphLaunch(exec, q));
        CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, q));
        CK(cudaGraphLaunch(exec, q));
        CK(cudaStreamSynchronize(q));
        cudaGraphExecDestroy(exec);
        s->launches = before + 2 * (s->launches - before);
        s->since_sort += n;
        s->n_substeps += n;
        for (int k = 0; k < n; k++) s->host_time += (double)dt;
    } else {
        for (int i = 0; i < n; i++) {
            s->g.ts = d + (size_t)i * per;
            launch_substep(s, a, q);
        }
        s->g.ts = nullptr;
    }
    s->profiling = prof;
    s->since_sort += n;
    s->n_substeps += n;
    s->canon_stale = true;
    for (int k = 0; k < n; k++) s->host_time += (double)dt;
    CK(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < h.size(); i += 2) t0 = std::min(t0, h[i]);
    for (size_t i = 0; i < h.size(); i += 2) {
        const bool ran = h[i] != ~0ull;
        out[i] = ran ? (long long)(h[i] - t0) : -1;
        out[i + 1] = ran ? (long long)(h[i + 1] - t0) : -1;
    }
    API_END(s)
}

int mpm_set_debug(MpmSolver* s, int on) {
    API_BEGIN(s)
    s->debug = on != 0;
    if (s->debug && !s->g.dbg_acc) {
        s->g.dbg_acc = s->dalloc<float4>((size_t)s->g.cap * BN);
    }
    if (!s->debug) s->g.dbg_acc = nullptr;
    API_END(s)
}

int mpm_export_grid(MpmSolver* s, float* grid_m, float* grid_v_in, float* grid_v_out, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    size_t n3 = (size_t)s->g.n * s->g.n * s->g.n;
    float *dm = nullptr, *dvi = nullptr, *dvo = nullptr;
    auto tmp = [&](size_t n) { void* p; CK(cudaMalloc(&p, n * sizeof(float))); CK(cudaMemsetAsync(p, 0, n * sizeof(float), q)); return (float*)p; };
    if (grid_m) dm = tmp(n3);
    if (grid_v_in) dvi = tmp(3 * n3);
    if (grid_v_out) dvo = tmp(3 * n3);
    k_export_grid<<<cdiv((long long)s->g.cap * BN, 256), 256, 0, q>>>(s->g, dm, dvi, dvo);
    if (dm) CK(cudaMemcpyAsync(grid_m, dm, n3 * sizeof(float), cudaMemcpyDefault, q));
    if (dvi) CK(cudaMemcpyAsync(grid_v_in, dvi, 3 * n3 * sizeof(float), cudaMemcpyDefault, q));
    if (dvo) CK(cudaMemcpyAsync(grid_v_out, dvo, 3 * n3 * sizeof(float), cudaMemcpyDefault, q));
    CK(cudaStreamSynchronize(q));
    if (dm) cudaFree(dm);
    if (dvi) cudaFree(dvi);
    if (dvo) cudaFree(dvo);
    API_END(s)
}

int mpm_set_profiling(MpmSolver* s, int on) {
    API_BEGIN(s)
    if (!on) drain_profile(s);
    s->profiling = on != 0;
    API_END(s)
}

int mpm_get_profile(MpmSolver* s, MpmProfile* out) {
    API_BEGIN(s)
    drain_profile(s);
    *out = s->prof;
    API_END(s)
}

int mpm_get_stats(MpmSolver* s, MpmStats* out, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (s->need_sort && s->have_state) resort(s, q);
    int n_slots = 0, flags[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(&n_slots, s->g.n_slots, sizeof(int), cudaMemcpyDeviceToHost, q));
    CK(cudaMemcpyAsync(flags, s->g.flags, sizeof flags, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    n_slots = std::min(n_slots, s->g.cap);
    unsigned long long* mask = nullptr;
    unsigned long long* cnt = nullptr;
    const size_t nmask = (size_t)s->g.cap + 1;  // one 64-bit word per block
    CK(cudaMalloc(&mask, nmask * sizeof(unsigned long long)));
    CK(cudaMalloc(&cnt, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(mask, 0, nmask * sizeof(unsigned long long), q));
    CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), q));
    if (s->Ne) k_mark_nodes<<<cdiv(s->Ne, 128), 128, 0, q>>>(s->g, s->Ne, (const float*)s->R.XE, 4, mask);
    if (s->Nt) k_mark_nodes<<<cdiv(s->Nt, 128), 128, 0, q>>>(s->g, s->Nt, s->R.TP, KP_F, mask);
    if (s->Nv) k_mark_nodes<<<cdiv(s->Nv, 128), 128, 0, q>>>(s->g, s->Nv, s->R.VP, VP_F, mask);
    k_popc<<<cdiv((long long)nmask, 256), 256, 0, q>>>(mask, (int)nmask, cnt);
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    cudaFree(mask);
    cudaFree(cnt);
    out->n_active_blocks = n_slots;
    out->n_active_nodes = (long long)h;
    out->n_resorts = s->n_resorts;
    out->n_substeps = s->n_substeps;
    out->overflow = (flags[0] ? 1 : 0) | (flags[1] ? 2 : 0);
    out->gpu_launches = s->launches;
    out->sim_time = s->host_time;
    API_END(s)
}

int mpm_debug_phase_clocks(MpmSolver* s, unsigned long long* out64, int reset) {
    API_BEGIN(s)
    CK(cudaDeviceSynchronize());
    if (out64) {
        std::vector<unsigned long long> h(64 * 64);
        CK(cudaMemcpy(h.data(), s->g.clk, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 64; i++) {
            out64[i] = 0;
            for (int c = 0; c < 64; c++) out64[i] += h[c * 64 + i];
        }
    }
    if (reset) CK(cudaMemset(s->g.clk, 0, 64 * 64 * sizeof(unsigned long long)));
    API_END(s)
}

int mpm_force_resort(MpmSolver* s) {
    API_BEGIN(s)
    s->since_sort = s->resort_interval;
    API_END(s)
}

}  // extern "C"

// ================================================================ caller-side glue on the device
#include "mpm_mesh.cuh"

#define STATELESS_BEGIN try {
#define STATELESS_END             \
    }                             \
    catch (const std::string& e) { \
        g_create_error = e;       \
        return -2;                \
    }                             \
    return 0;

extern "C" {

int mpm_cloth_normalisation(const float* verts_wld, int n_verts, float* scale_shift4, void* stream) {
    STATELESS_BEGIN
    cudaStream_t q = (cudaStream_t)stream;
    if (!verts_wld || n_verts <= 0 || !scale_shift4) throw std::string("mpm_cloth_normalisation: bad arguments");
    float* d = nullptr;
    CK(cudaMalloc(&d, 6 * sizeof(float)));
    const float init[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    CK(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, q));
    k_minmax<<<std::min(cdiv(n_verts, 256), 592), 256, 0, q>>>(verts_wld, n_verts, d);
    float h[6];
    CK(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    cudaFree(d);
    // train_material_params.py:365-370 in fp32, as torch evaluates it
    float max_diff = 0.f;
    for (int a = 0; a < 3; a++) max_diff = std::max(max_diff, h[3 + a] - h[a]);
    const float scale = 1.0f / max_diff;
    scale_shift4[0] = scale;
    for (int a = 0; a < 3; a++) scale_shift4[1 + a] = 1.0f - ((h[a] + h[3 + a]) / 2.0f) * scale;
    STATELESS_END
}

int mpm_build_cloth_particles(const float* verts_wld, const int* faces, int n_verts, int n_faces, float thickness, float scale,
                              const float shift[3], const MpmClothParticles* out, void* stream) {
    STATELESS_BEGIN
    cudaStream_t q = (cudaStream_t)stream;
    if (!verts_wld || !faces || !out || n_verts <= 0 || n_faces <= 0) throw std::string("mpm_build_cloth_particles: bad arguments");
    if (!out->x) throw std::string("mpm_build_cloth_particles: out->x is required (the sim-space vertices live in its tail)");
    float* verts_sim = out->x + 3 * (size_t)n_faces;  // canonical order [elements | vertices] (train_material_params.py:387)
    k_wld2sim<<<cdiv(n_verts, 256), 256, 0, q>>>(verts_wld, n_verts, scale, shift[0], shift[1], shift[2], verts_sim);
    if (out->vol) CK(cudaMemsetAsync(out->vol + n_faces, 0, (size_t)n_verts * sizeof(float), q));
    ClothOut o{out->x, out->init_dir, out->rest_dir, out->rest_dir_inv, out->vol, out->vol ? out->vol + n_faces : nullptr};
    k_cloth_faces<<<cdiv(n_faces, 128), 128, 0, q>>>(verts_sim, faces, n_faces, thickness, o);
    CK(cudaGetLastError());
    STATELESS_END
}

int mpm_export_cloth_verts(MpmSolver* s, float scale, const float shift[3], float* out_wld, const long long* scatter_idx,
                           float* full_verts, const float* target, double* sse, void* stream) {
    API_BEGIN(s)
    cudaStream_t q = (cudaStream_t)stream;
    if (!s->have_state) throw std::string("mpm_export_cloth_verts before mpm_import_state");
    if (s->need_sort) resort(s, q);
    if (sse) CK(cudaMemsetAsync(sse, 0, sizeof(double), q));
    if (s->Nv)
        k_export_verts<<<cdiv(s->Nv, 256), 256, 0, q>>>(s->Nv, s->R.VP, s->R.permV, scale, shift[0], shift[1], shift[2], out_wld, scatter_idx,
                                                       full_verts, target, sse);
    s->launches++;
    CK(cudaGetLastError());
    API_END(s)
}

int mpm_face_frames(const float* verts, const int* faces, int n_faces, float* center, float* orien, float* quat, float* scale, void* stream) {
    STATELESS_BEGIN
    if (!verts || !faces || n_faces <= 0) throw std::string("mpm_face_frames: bad arguments");
    k_face_frames<<<cdiv(n_faces, 128), 128, 0, (cudaStream_t)stream>>>(verts, faces, n_faces, FaceFrames{center, orien, quat, scale});
    CK(cudaGetLastError());
    STATELESS_END
}

// "v x y z\n" per vertex with the shortest decimal that reads back as the same float32 (what the reference's f-string of a
// numpy float32 prints), then `tail` (the texture-coordinate / face lines) verbatim; one buffered write
int mpm_write_obj(const char* path, const float* verts_host, int n_verts, const char* tail, long long tail_len) {
    STATELESS_BEGIN
    if (!path || (!verts_host && n_verts > 0)) throw std::string("mpm_write_obj: bad arguments");
    std::string buf;
    buf.resize((size_t)n_verts * 52 + 16);
    char* p = &buf[0];
    auto put = [&](float v) {
        char* b = p;
        auto r = std::to_chars(p, p + 24, v);
        p = r.ptr;
        bool plain = true;  // "1" -> "1.0" as Python prints floats
        for (char* c = b; c < p; c++) if (*c == '.' || *c == 'e' || *c == 'n' || *c == 'i') plain = false;
        if (plain) { *p++ = '.'; *p++ = '0'; }
    };
    for (int i = 0; i < n_verts; i++) {
        *p++ = 'v';
        for (int a = 0; a < 3; a++) { *p++ = ' '; put(verts_host[3 * (size_t)i + a]); }
        *p++ = '\n';
    }
    FILE* f = fopen(path, "wb");
    if (!f) throw std::string("mpm_write_obj: cannot open ") + path;
    bool ok = fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) == (size_t)(p - buf.data());
    if (ok && tail && tail_len > 0) ok = fwrite(tail, 1, (size_t)tail_len, f) == (size_t)tail_len;
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::string("mpm_write_obj: write failed for ") + path;
    STATELESS_END
}

}  // extern "C"
