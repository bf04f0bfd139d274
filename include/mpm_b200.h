/*
 * mpm_b200.h -- C-ABI of the B200-native MPM substep solver (libmpm_b200.so).
 *
 * The reference (KAISTChangmin/MPMAvatar) has no FFI boundary: its boundary is the
 * in-process Python API of warp_mpm.  Each entry point below
 * names the reference interface it stands behind; mpmavatar_b200/warp_mpm/ binds
 * them with ctypes and re-exposes the reference's own class / method names.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / warp types.
 *  - every data pointer may be a DEVICE or a (pinned or pageable) HOST pointer:
 *    copies use cudaMemcpyDefault (UVA).  Device pointers are the fast path.
 *  - all work is enqueued on the caller's stream (cudaStream_t passed as void*);
 *    no entry point synchronises the host unless its comment says so.
 *  - one handle per GPU, not thread-safe (same as the reference: single Python thread).
 *  - return value 0 = ok, <0 = error; mpm_last_error() returns the message.
 *  - "canonical" arrays are the reference's layouts: particle order
 *    [elements | traditional | vertices] (train_material_params.py:387), row-major
 *    mat33, faces as float vertex-local indices (mpm_data_structure.py:41,211-215).
 */
#ifndef MPM_B200_H
#define MPM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MpmSolver MpmSolver;

/* MPMWARP.__init__/initialize (warp_mpm/mpm_solver.py:14-51) +
 * MPMStateStruct.init/init_grid (mpm_data_structure.py:51-156) +
 * MPMModelStruct.init_other_params (mpm_data_structure.py:686-715) */
typedef struct {
    int n_particles, n_elements, n_vertices; /* n_traditional = N - Ne - Nv */
    int n_grid;
    float grid_lim;
    int n_mesh_v, n_mesh_f; /* body mesh sizes, 0 = no mesh */
    int num_joint_v, num_joint_f;
    int device;           /* CUDA device ordinal */
    int resort_interval;  /* substeps between particle re-sorts; <=0 = default */
} MpmConfig;

/* model scalars: MPMModelStruct fields written by set_parameters_dict
 * (mpm_solver.py:57-126) and init_other_params */
typedef struct {
    int material;  /* 0 jelly 1 metal 2 sand 3 foam 4 snow 5 plasticine 6 neo-hookean 7 cloth */
    int hardening; /* model.hardening == 1 enables hardening (mpm_utils.py:249) */
    float friction_coeff, alpha;
    float g[3];
    float rpic_damping, grid_v_damping_scale;
    float xi, plastic_viscosity, softening;
} MpmModelParams;

/* canonical particle arrays; NULL members are skipped.
 * MPMStateStruct.from_torch / reset_state / continue_from_torch / reset_density
 * (mpm_data_structure.py:158-467), MPMWARP.set_E_nu / prepare_mu_lam
 * (mpm_solver.py:128-227) */
typedef struct {
    float *x, *v;           /* [N,3] */
    float *C;               /* [N,9] */
    float *F, *F_trial;     /* [Nnv,9]; only the traditional rows are used */
    float *stress;          /* [Nnv,9] (export only) */
    float *d;               /* [Ne,9] */
    float *R_inv;           /* [Ne,3] */
    float *faces;           /* [Ne,3] float vertex-local indices */
    float *vertex_force;    /* [Nv,3] (export only) */
    float *vol, *mass;      /* [N] */
    float *mu, *lam, *gamma, *kappa, *yield_stress; /* [N] */
} MpmParticleArrays;

/* per-call inputs of MPMWARP.p2g2p (mpm_solver.py:229-231): body mesh points /
 * velocities and prescribed joint velocities.  NULL = not given. */
typedef struct {
    const float *mesh_x, *mesh_v;             /* [n_mesh_v,3] */
    const float *joint_traditional_v;         /* [n_joint_t,3]: pins the LAST n_joint_t traditional particles */
    int n_joint_t;
    const float *joint_verts_v, *joint_faces_v; /* [num_joint_v,3], [num_joint_f,3] */
    int device_inputs; /* 1: every non-NULL pointer above is a DEVICE pointer -- the inputs are then staged by ONE kernel
                        * instead of one cudaMemcpyAsync each, and a single substep replays a captured graph: two driver
                        * calls per p2g2p (the unchanged callers make one call per substep, train_material_params.py:622-626) */
} MpmFrameInputs;

typedef struct {
    int n_active_blocks;       /* allocated 4^3-node grid blocks */
    long long n_active_nodes;  /* distinct nodes in the union of all particle stencils */
    int n_resorts;
    long long n_substeps;
    int overflow;              /* 1 = block pool exhausted or a scatter hit an unallocated block */
    int gpu_launches;          /* kernels launched by the library since creation */
    double sim_time;
} MpmStats;

/* cumulative per-phase device time in ms (filled only while profiling is on);
 * stands behind MPMWARP.time_profile / print_time_profile (mpm_solver.py:16,538-541) */
typedef struct {
    float stress_ms, p2g_ms, collider_scatter_ms, mover_scatter_ms, grid_ms, g2p_v_ms, g2p_e_ms, resort_ms;
    long long n_substeps;
} MpmProfile;

int mpm_create(const MpmConfig *cfg, MpmSolver **out);
void mpm_destroy(MpmSolver *s);
const char *mpm_last_error(MpmSolver *s); /* s may be NULL: last creation error */

int mpm_set_model(MpmSolver *s, const MpmModelParams *p);

/* copy the non-NULL canonical arrays into the solver (async on stream); positions
 * trigger a particle re-sort + sparse-grid rebuild at the next step */
int mpm_import_state(MpmSolver *s, const MpmParticleArrays *a, void *stream);
/* write the current state back to the non-NULL canonical arrays in ORIGINAL particle
 * order (wp.to_torch(state.particle_x), train_material_params.py:628) */
int mpm_export_state(MpmSolver *s, const MpmParticleArrays *a, void *stream);

/* body-mesh topology, MPMWARP.initialize (mpm_solver.py:45-51): faces [n_mesh_f,3] int32 */
int mpm_set_body_mesh(MpmSolver *s, const int *faces, const float *points0, void *stream);
/* add_mesh_collider (mpm_solver.py:805-919) */
int mpm_add_mesh_collider(MpmSolver *s, float friction);
/* add_particle_mover (mpm_solver.py:661-802) */
int mpm_add_particle_mover(MpmSolver *s);
/* add_surface_collider (mpm_solver.py:564-658); surface_type 0 sticky, 1 slip, 11 cut, 2 other */
int mpm_add_surface_collider(MpmSolver *s, const float point[3], const float normal[3], int surface_type,
                             float friction, float start_time, float end_time);
/* set_velocity_on_cuboid (mpm_solver.py:929-984) */
int mpm_set_velocity_on_cuboid(MpmSolver *s, const float point[3], const float size[3], const float velocity[3],
                               float start_time, float end_time, int reset);
/* add_bounding_box (mpm_solver.py:986-1053) */
int mpm_add_bounding_box(MpmSolver *s, float start_time, float end_time);
/* enforce_grid_velocity_by_mask (mpm_solver.py:1330-1355); mask [n_grid^3] int32 */
int mpm_enforce_grid_velocity_by_mask(MpmSolver *s, const int *mask, void *stream);
/* pre-P2G particle operations (mpm_solver.py:1058-1328, 1360-1417).  mask [N] int32 in canonical order.
 * kind 0: v += force/mass*dt where mask==1 (add_impulse_on_particles)
 * kind 1: v += force*dt where mask>=1      (add_impulse_on_particles_with_mask)
 * kind 2: v = velocity where mask==1        (enforce_particle_velocity_translation / _by_mask)
 * Applied as the reference does: all impulses (kinds 0, 1) in the order added, then all velocity modifiers. */
int mpm_add_particle_op(MpmSolver *s, int kind, const float vec[3], const int *mask, float start_time,
                        float end_time, void *stream);
/* enforce_particle_velocity_rotation (mpm_solver.py:1156-1256): particles with mask==1 (the cylinder selection made by
 * the caller at call time) get v = -h sin(t) rot * axis1 + h cos(t) rot * axis2 + translation_scale * normal, with
 * h, t the polar coordinates of the particle about (point, normal); normal is unit, axis1/axis2 as the reference builds them */
int mpm_add_particle_rotation(MpmSolver *s, const float point[3], const float normal[3], const float axis1[3],
                              const float axis2[3], float rotation_scale, float translation_scale, const int *mask,
                              float start_time, float end_time, void *stream);

/* nsub substeps of MPMWARP.p2g2p (mpm_solver.py:229-536).  Substep k uses body points
 * mesh_x + (float)(dt*k) * mesh_v, which is what the callers compute on the host side
 * (train_material_params.py:622-626); nsub = 1 is exactly one p2g2p call. */
int mpm_step(MpmSolver *s, float dt, int nsub, const MpmFrameInputs *in, void *stream);

/* ---- sharded runs (the reference is single-GPU, so these have no reference counterpart).
 * One p2g2p substep split around the grid: mpm_step_scatter = constitutive update + P2G + body / joint
 * scatters, mpm_step_gather = grid update + G2P.  Between the two the caller sum-reduces the packed
 * buffer of the grid blocks shared with other ranks (one all-reduce per substep):
 *   mpm_shared_pack -> all-reduce(buf, n_shared*64*8 floats) -> mpm_shared_unpack.
 * Ghost copies of vertices owned by another rank are ordinary vertex particles with mass 0: their
 * scatter is then exactly the force term dt*w*f, which is linear, so partial vertex forces need no
 * exchange of their own.  mpm_get_active_blocks synchronises; coords are bx | by<<10 | bz<<20. */
int mpm_step_scatter(MpmSolver *s, float dt, const MpmFrameInputs *in, void *stream);
/* The same, nsub substeps driven from C.  `exchange` must enqueue a sum-reduction of buf[0..n_floats) over
 * the ranks on `stream` (the caller owns the communicator); `rebuild` must call mpm_get_potential_blocks,
 * agree on the shared list with the other ranks and call mpm_set_shared_blocks; it is invoked every
 * `refresh` substeps.  buf must hold 512 floats per shared block.  Both return 0 on success. */
typedef int (*MpmExchangeFn)(void *ctx, float *buf, int n_floats);
typedef int (*MpmRebuildFn)(void *ctx);
int mpm_step_sharded(MpmSolver *s, float dt, int nsub, const MpmFrameInputs *in, float *buf, int refresh,
                     MpmExchangeFn exchange, MpmRebuildFn rebuild, void *ctx, void *stream);
/* The same again with the exchange done inside the library: mpm_attach_comm gives the solver its own NCCL
 * communicator (id = ncclUniqueId bytes from mpm_comm_unique_id on rank 0, distributed by the caller; libnccl.so.2
 * is dlopen'ed, inside a torch process that is torch's own copy), and mpm_step_sharded_nccl replays captured
 * windows of 8 substeps -- kernels AND the ncclAllReduce of the shared blocks in one CUDA graph -- so that no
 * host code runs between substeps.  Every `refresh` substeps the shared list is rebuilt on the stream as well
 * (blocks reachable within `margin` cells are marked, the marks byte-summed over the ranks, blocks with count >= 2
 * compacted in ascending order); only buffer growth synchronises.  mpm_shared_info synchronises. */
int mpm_comm_unique_id(char *out128);
int mpm_attach_comm(MpmSolver *s, const char *id128, int rank, int nranks);
/* The same stepping over a transport the CALLER owns instead of NCCL: `allgather` must gather `nbytes` host bytes from
 * every rank into recv[nranks * nbytes] in rank order and block until done (torch.distributed / gloo, MPI, ...).  Only
 * the set-up traffic uses it -- the CUDA-IPC handles of the receive areas and the block marks of a shared-list rebuild;
 * the per-substep exchange stays peer-to-peer, fused into the grid update (mode 2), also
 * between ranks that share one GPU (which NCCL refuses), so the kernels can be verified on a single-GPU box.  Where no peer mapping is possible the exchange
 * falls back to pack -> host all-gather -> rank-ordered sum -> unpack (mode 3; synchronises every substep). */
typedef int (*MpmHostAllGatherFn)(void *ctx, const void *send, void *recv, int nbytes);
int mpm_attach_host_comm(MpmSolver *s, int rank, int nranks, MpmHostAllGatherFn allgather, void *ctx);
int mpm_step_sharded_nccl(MpmSolver *s, float dt, int nsub, const MpmFrameInputs *in, int refresh, int margin, void *stream);
int mpm_shared_info(MpmSolver *s, int *n_shared, int *cap_blocks, int *n_rebuilds, void *stream);
/* how the shared blocks travel: 0 caller's collective (callbacks), 1 ncclAllReduce inside the captured windows,
 * 3 all-reduce through the caller's host all-gather, 2 peer-to-peer: with <= 8 ranks on one NVLink domain every rank
 * maps every peer's receive area (CUDA IPC) and the grid update pushes / pulls the parts itself (MPM_B200_P2P=0 forces 1) */
int mpm_shared_mode(MpmSolver *s);
int mpm_step_gather(MpmSolver *s, float dt, void *stream);
int mpm_get_active_blocks(MpmSolver *s, int *coords, int cap, int *n, void *stream);
/* blocks this rank can activate while its particles move at most `margin` cells (synchronises) */
int mpm_get_potential_blocks(MpmSolver *s, int margin, int *coords, int cap, int *n, void *stream);
int mpm_set_shared_blocks(MpmSolver *s, const int *coords, int n, void *stream);
int mpm_shared_pack(MpmSolver *s, float *buf, void *stream);
int mpm_shared_unpack(MpmSolver *s, const float *buf, void *stream);

/* self.time (mpm_solver.py:28,536) */
int mpm_set_time(MpmSolver *s, double t);

/* test / measurement hooks */
/* dense [n^3] grid_m, [n^3,3] grid_v_in, [n^3,3] grid_v_out of the LAST substep (NULL skipped);
 * requires mpm_set_debug(s, 1) before the step.  Synchronises. */
int mpm_export_grid(MpmSolver *s, float *grid_m, float *grid_v_in, float *grid_v_out, void *stream);
int mpm_set_debug(MpmSolver *s, int on);
int mpm_set_profiling(MpmSolver *s, int on); /* per-phase CUDA events; disables graph replay */
int mpm_get_profile(MpmSolver *s, MpmProfile *out);   /* synchronises */
int mpm_get_stats(MpmSolver *s, MpmStats *out, void *stream); /* synchronises */
int mpm_force_resort(MpmSolver *s);
/* timeline probe: n <= 64 substeps with per-kernel first-start / last-end stamps of the GPU global timer;
 * out[n][8][2] ns relative to the first stamp (-1: kernel not launched).  Kernel ids: 0 P2G elements (+ cloth stress),
 * 1 P2G traditional, 2 P2G vertices, 3 body/joint scatter, 4 grid update, 5 G2P vertices, 6 G2P traditional,
 * 7 G2P elements.  The kernels overlap under programmatic dependent launch, which CUDA events cannot resolve. */
int mpm_measure_timeline(MpmSolver *s, float dt, int n, const MpmFrameInputs *in, long long *out, void *stream);
/* the same for the SHARDED chain (collective; needs the peer-to-peer exchange, i.e. at least one mpm_step_sharded_nccl call
 * before): n <= 32, out[n][10][2]; ids 0-7 as above; 8 = the push phase at the head of the grid update (flagged stores of
 * this rank's parts into the members' receive areas), 9 = its second pass: the nodes of shared blocks, including the wait
 * for their members' parts (the neighbours' skew).  The whole exchange is fused into k_grid_update<true>. */
int mpm_measure_timeline_sharded(MpmSolver *s, float dt, int n, const MpmFrameInputs *in, long long *out, void *stream);
/* latency analysis (builds with -DMPM_CLK only fill it): 8 kernels x 8 per-warp phase cycle sums; synchronises */
int mpm_debug_phase_clocks(MpmSolver *s, unsigned long long *out64, int reset);

/* ---- caller-side glue around the hot path, on the device.  The first, second, fourth and
 * fifth are stateless (no solver handle; errors through mpm_last_error(NULL)); device pointers unless said otherwise. */

/* wld2sim normalisation of setup_simulation (train_material_params.py:365-373): scale_shift4 (HOST) = {scale, shift[3]} with
 * scale = 1 / max extent of verts_wld, shift = (1,1,1) - box centre * scale.  Synchronises. */
int mpm_cloth_normalisation(const float *verts_wld, int n_verts, float *scale_shift4, void *stream);

/* Particle construction from the tracked garment mesh: verts_sim = verts_wld * scale + shift, then compute_dir_vol and
 * compute_rest_dir_inv (train_material_params.py:508-515, 533-553) per face, element centroids (:379).  NULL outputs
 * (except x) are skipped. */
typedef struct {
    float *x;            /* [n_faces + n_verts, 3] canonical positions: element centroids, then the sim-space vertices */
    float *vol;          /* [n_faces + n_verts] element_vol = 0.25 * thickness * area, then vertex_vol = sum over incident faces */
    float *init_dir;     /* [n_faces, 9] row-major d = [d1 d2 normalize(d1 x d2)] as columns */
    float *rest_dir;     /* [n_faces, 3] R11 R12 R22 of the QR of [d1 d2] */
    float *rest_dir_inv; /* [n_faces, 3] 1/R11, -R12/(R11 R22), 1/R22 */
} MpmClothParticles;
int mpm_build_cloth_particles(const float *verts_wld, const int *faces, int n_verts, int n_faces, float thickness, float scale,
                              const float shift[3], const MpmClothParticles *out, void *stream);

/* Cloth vertex positions straight from the solver's sorted records: original vertex order, world coordinates
 * ((p - shift) / scale: sim2wld, train_material_params.py:373,630,812).  Optional: out_wld [n_vertices,3]; scatter into a
 * full-body vertex array full_verts at scatter_idx [n_vertices] (int64, reordered_cloth_v_idx, :814); sum of squared
 * differences to target [n_vertices,3] into *sse (device double; F.mse_loss = sse / (3 n_vertices), :631). */
int mpm_export_cloth_verts(MpmSolver *s, float scale, const float shift[3], float *out_wld, const long long *scatter_idx,
                           float *full_verts, const float *target, double *sse, void *stream);

/* Per-frame OBJ (train_material_params.py:819-821): "v x y z" lines from HOST float32 vertices -- every number the shortest
 * decimal that reads back as the same float32 -- followed by `tail` (the vt / f lines) verbatim. */
int mpm_write_obj(const char *path, const float *verts_host, int n_verts, const char *tail, long long tail_len);

/* Hand-off to the renderer without the OBJ detour: what MeshGaussianModel.set_mesh_by_verts computes from the simulated
 * vertices (scene/mesh_gaussian_model.py:137-146, utils/graphics_utils.py:88-107): face centre [F,3], orientation matrix
 * [F,9] (columns a0 a1 a2), unit quaternion wxyz [F,4] (defined up to sign), scale [F].  NULL outputs are skipped. */
int mpm_face_frames(const float *verts, const int *faces, int n_faces, float *center, float *orien, float *quat, float *scale,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif
