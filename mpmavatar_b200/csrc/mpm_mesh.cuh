// Caller-side glue around the hot path, on the device:
//   * particle construction from the tracked garment mesh: wld2sim normalisation, compute_dir_vol,
//     compute_rest_dir_inv(_from_vf)               (train_material_params.py:365-373, 508-553)
//   * export: vertex positions straight from the solver's sorted records, un-permuted, sim2wld applied, optionally
//     scattered into the full-body vertex array and compared with the tracked frame (MSE)
//                                                  (train_material_params.py:628-631, 811-817)
//   * OBJ text: shortest round-trip decimal of every float32 (binary-safe), one buffered write
//                                                  (train_material_params.py:819-821)
//   * hand-off to the renderer: per-face frame (orientation matrix, quaternion, scale, centre) of the mesh-bound
//     Gaussians from the simulated vertices         (scene/mesh_gaussian_model.py:137-146, utils/graphics_utils.py)
// Included at the end of mpm_b200.cu (one translation unit, one library).
#pragma once
#include <charconv>
#include <cstdio>

namespace mpm {

// ---- wld2sim normalisation: scale = 1 / max extent, shift = (1,1,1) - centre * scale (train_material_params.py:365-373)
__global__ void k_minmax(const float* __restrict__ v, int n, float* __restrict__ out6) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int a = 0; a < 3; a++) { float x = v[3 * (size_t)i + a]; lo[a] = fminf(lo[a], x); hi[a] = fmaxf(hi[a], x); }
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {  // float min / max through the ordered-int trick
            int l = __float_as_int(lo[a]), h = __float_as_int(hi[a]);
            if (l >= 0) atomicMin(reinterpret_cast<int*>(out6) + a, l); else atomicMax(reinterpret_cast<unsigned*>(out6) + a, (unsigned)l);
            if (h >= 0) atomicMax(reinterpret_cast<int*>(out6) + 3 + a, h); else atomicMin(reinterpret_cast<unsigned*>(out6) + 3 + a, (unsigned)h);
        }
    }
}

// ---- compute_dir_vol + compute_rest_dir_inv (train_material_params.py:508-515, 533-553), one thread per face.
// verts_sim [Nv,3] = wld2sim(verts) is written by k_wld2sim first; element_vol is scattered to the three corners with
// float atomics (the reference's index_add_ is an atomic scatter as well).
__global__ void k_wld2sim(const float* __restrict__ vw, int n, float scale, float sx, float sy, float sz, float* __restrict__ vs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vs[3 * (size_t)i] = vw[3 * (size_t)i] * scale + sx;
    vs[3 * (size_t)i + 1] = vw[3 * (size_t)i + 1] * scale + sy;
    vs[3 * (size_t)i + 2] = vw[3 * (size_t)i + 2] * scale + sz;
}
struct ClothOut {
    float* centroid;   // [Ne,3] mean of the three corners (train_material_params.py:379)
    float* init_dir;   // [Ne,9] row-major, columns d1 d2 d3
    float* rest_dir;   // [Ne,3] R11 R12 R22
    float* rest_inv;   // [Ne,3] iR11 iR12 iR22
    float* elem_vol;   // [Ne]
    float* vert_vol;   // [Nv], zeroed by the caller
};
__global__ void k_cloth_faces(const float* __restrict__ v, const int* __restrict__ faces, int Ne, float thickness, ClothOut o) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Ne) return;
    const int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
    float p0[3], p1[3], p2[3], d1[3], d2[3];
    for (int a = 0; a < 3; a++) { p0[a] = v[3 * (size_t)i0 + a]; p1[a] = v[3 * (size_t)i1 + a]; p2[a] = v[3 * (size_t)i2 + a]; }
    for (int a = 0; a < 3; a++) { d1[a] = __fsub_rn(p1[a], p0[a]); d2[a] = __fsub_rn(p2[a], p0[a]); }
    // torch evaluates these as separate fp32 kernels: no contraction across them
    float c[3] = {__fsub_rn(__fmul_rn(d1[1], d2[2]), __fmul_rn(d1[2], d2[1])), __fsub_rn(__fmul_rn(d1[2], d2[0]), __fmul_rn(d1[0], d2[2])),
                  __fsub_rn(__fmul_rn(d1[0], d2[1]), __fmul_rn(d1[1], d2[0]))};
    const float cn = len3_rn(c);
    if (o.init_dir) {
        float* D = o.init_dir + 9 * (size_t)f;
        for (int a = 0; a < 3; a++) { D[3 * a] = d1[a]; D[3 * a + 1] = d2[a]; D[3 * a + 2] = __fdiv_rn(c[a], cn); }
    }
    const float R11 = len3_rn(d1);
    const float R12 = __fdiv_rn(dot3_rn(d1, d2), R11);
    const float q = __fdiv_rn(R12, R11);
    const float u[3] = {__fsub_rn(d2[0], __fmul_rn(q, d1[0])), __fsub_rn(d2[1], __fmul_rn(q, d1[1])), __fsub_rn(d2[2], __fmul_rn(q, d1[2]))};
    const float R22 = len3_rn(u);
    if (o.rest_dir) { o.rest_dir[3 * (size_t)f] = R11; o.rest_dir[3 * (size_t)f + 1] = R12; o.rest_dir[3 * (size_t)f + 2] = R22; }
    if (o.rest_inv) {
        const float iR11 = __fdiv_rn(1.0f, R11), iR22 = __fdiv_rn(1.0f, R22);
        o.rest_inv[3 * (size_t)f] = iR11;
        o.rest_inv[3 * (size_t)f + 1] = __fmul_rn(__fmul_rn(-R12, iR11), iR22);
        o.rest_inv[3 * (size_t)f + 2] = iR22;
    }
    const float ev = __fmul_rn(__fmul_rn(0.25f, thickness), __fmul_rn(0.5f, cn));
    if (o.elem_vol) o.elem_vol[f] = ev;
    if (o.vert_vol) { atomicAdd(&o.vert_vol[i0], ev); atomicAdd(&o.vert_vol[i1], ev); atomicAdd(&o.vert_vol[i2], ev); }
    if (o.centroid)
        for (int a = 0; a < 3; a++) o.centroid[3 * (size_t)f + a] = __fdiv_rn(__fadd_rn(__fadd_rn(p0[a], p1[a]), p2[a]), 3.0f);
}

// ---- export: cloth vertex positions from the SORTED vertex records, in original vertex order, in world coordinates
// ((p - shift) / scale, train_material_params.py:373,630); optionally written into a full-body vertex array at
// scatter_idx (:814) and compared with a target (squared-error sum for F.mse_loss, :631)
__global__ void k_export_verts(int Nv, const float* __restrict__ VP, const uint32_t* __restrict__ permV, float scale, float sx, float sy, float sz,
                               float* __restrict__ out, const long long* __restrict__ scatter_idx, float* __restrict__ full, const float* __restrict__ target,
                               double* __restrict__ sse) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < Nv) {
        const float* p = VP + (size_t)i * VP_F;
        const int c = (int)permV[i];
        const float w[3] = {__fdiv_rn(__fsub_rn(p[0], sx), scale), __fdiv_rn(__fsub_rn(p[1], sy), scale), __fdiv_rn(__fsub_rn(p[2], sz), scale)};
        if (out) for (int a = 0; a < 3; a++) out[3 * (size_t)c + a] = w[a];
        if (full && scatter_idx) { const long long j = scatter_idx[c]; for (int a = 0; a < 3; a++) full[3 * (size_t)j + a] = w[a]; }
        if (target) for (int a = 0; a < 3; a++) { const double d = (double)w[a] - (double)target[3 * (size_t)c + a]; e += d * d; }
    }
    if (sse) {
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(sse, e);
    }
}

// ---- renderer hand-off: per-face frame of the mesh-bound Gaussians (scene/mesh_gaussian_model.py:137-146 with
// utils/graphics_utils.py compute_face_orientation): a0 = normalize(v1 - v0), a1 = normalize((v1 - v0) x (v2 - v0)),
// a2 = -normalize(a1 x a0); orientation = [a0 a1 a2] as COLUMNS; scale = (|v1 - v0| + |a2 . (v2 - v0)|) / 2;
// centre = face mean; quaternion (w,x,y,z) of the orientation matrix.
__device__ __forceinline__ void rotmat_to_quat_wxyz(const float (&m)[3][3], float* q) {
    // the branch on the largest diagonal term every robust conversion uses (roma.rotmat_to_unitquat gives xyzw; the
    // reference reorders to wxyz, mesh_gaussian_model.py:144)
    const float t = m[0][0] + m[1][1] + m[2][2];
    float w, x, y, z;
    if (t > 0.0f) { float s = sqrtf(t + 1.0f) * 2.0f; w = 0.25f * s; x = (m[2][1] - m[1][2]) / s; y = (m[0][2] - m[2][0]) / s; z = (m[1][0] - m[0][1]) / s; }
    else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) { float s = sqrtf(1.0f + m[0][0] - m[1][1] - m[2][2]) * 2.0f; w = (m[2][1] - m[1][2]) / s; x = 0.25f * s; y = (m[0][1] + m[1][0]) / s; z = (m[0][2] + m[2][0]) / s; }
    else if (m[1][1] > m[2][2]) { float s = sqrtf(1.0f + m[1][1] - m[0][0] - m[2][2]) * 2.0f; w = (m[0][2] - m[2][0]) / s; x = (m[0][1] + m[1][0]) / s; y = 0.25f * s; z = (m[1][2] + m[2][1]) / s; }
    else { float s = sqrtf(1.0f + m[2][2] - m[0][0] - m[1][1]) * 2.0f; w = (m[1][0] - m[0][1]) / s; x = (m[0][2] + m[2][0]) / s; y = (m[1][2] + m[2][1]) / s; z = 0.25f * s; }
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
struct FaceFrames {
    float* center;  // [F,3]
    float* orien;   // [F,9] row-major
    float* quat;    // [F,4] wxyz
    float* scale;   // [F]
};
__global__ void k_face_frames(const float* __restrict__ verts, const int* __restrict__ faces, int nF, FaceFrames o) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
    float v0[3], v1[3], v2[3], e1[3], e2[3];
    for (int a = 0; a < 3; a++) { v0[a] = verts[3 * (size_t)i0 + a]; v1[a] = verts[3 * (size_t)i1 + a]; v2[a] = verts[3 * (size_t)i2 + a]; }
    for (int a = 0; a < 3; a++) { e1[a] = v1[a] - v0[a]; e2[a] = v2[a] - v0[a]; }
    const float eps = 1e-10f;  // safe_normalize floor (graphics_utils)
    const float s0 = len3(e1[0], e1[1], e1[2]);
    float a0[3], a1[3], a2[3];
    for (int a = 0; a < 3; a++) a0[a] = e1[a] / fmaxf(s0, eps);
    float c[3] = {a0[1] * e2[2] - a0[2] * e2[1], a0[2] * e2[0] - a0[0] * e2[2], a0[0] * e2[1] - a0[1] * e2[0]};
    const float cn = len3(c[0], c[1], c[2]);
    for (int a = 0; a < 3; a++) a1[a] = c[a] / fmaxf(cn, eps);
    float d[3] = {a1[1] * a0[2] - a1[2] * a0[1], a1[2] * a0[0] - a1[0] * a0[2], a1[0] * a0[1] - a1[1] * a0[0]};
    const float dn = len3(d[0], d[1], d[2]);
    for (int a = 0; a < 3; a++) a2[a] = -d[a] / fmaxf(dn, eps);
    const float s1 = fabsf(a2[0] * e2[0] + a2[1] * e2[1] + a2[2] * e2[2]);
    if (o.center) for (int a = 0; a < 3; a++) o.center[3 * (size_t)f + a] = (v0[a] + v1[a] + v2[a]) / 3.0f;
    const float m[3][3] = {{a0[0], a1[0], a2[0]}, {a0[1], a1[1], a2[1]}, {a0[2], a1[2], a2[2]}};
    if (o.orien) for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) o.orien[9 * (size_t)f + 3 * r + cc] = m[r][cc];
    if (o.quat) rotmat_to_quat_wxyz(m, o.quat + 4 * (size_t)f);
    if (o.scale) o.scale[f] = (fmaxf(s0, eps) + s1) / 2.0f;  // length() clamps like safe_normalize does
}

}  // namespace mpm
