/* titgpu — C ABI of the B200-native WCSPH particle step.
 *
 * The reference (Jhuighuy/TitSolver) has no plugin / FFI layer for this path:
 * it is header-only C++ templates instantiated in
 * /root/reference/source/titwcsph/wcsph.cpp. This C ABI sits *underneath* a
 * C++ facade that keeps the reference's template surface (include/tit/...);
 * each entry point names the reference interface it replaces.
 *
 * Conventions: opaque handle, `int` status (0 = ok), no exceptions cross the
 * boundary, one host thread per context, one context per GPU. All particle
 * arrays are host pointers in the reference's particle order (fluid particles
 * first, then fixed ones; /root/reference/source/tit/sph/particle_array.hpp:
 * 188-199, 233-239). `stride_bytes` is the distance between consecutive
 * particles' values (0 = packed), so that the reference's padded 32-byte
 * Vec<double,3> arrays can be passed directly.
 *
 * There is no CPU fallback: every call fails with a CUDA error if no device
 * is present.
 */
#ifndef TITGPU_H
#define TITGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TITGPU_API __attribute__((visibility("default")))
#else
#define TITGPU_API
#endif

typedef struct titgpu_ctx titgpu_ctx;

/* Smoothing kernels (/root/reference/source/tit/sph/kernel.hpp:428-454). */
enum { TITGPU_KERNEL_CUBIC_SPLINE = 0, TITGPU_KERNEL_QUARTIC_SPLINE = 1, TITGPU_KERNEL_QUINTIC_SPLINE = 2,
       TITGPU_KERNEL_QUARTIC_WENDLAND = 3, TITGPU_KERNEL_SIXTH_ORDER_WENDLAND = 4, TITGPU_KERNEL_EIGHTH_ORDER_WENDLAND = 5 };
/* Equations of state (sph/equation_of_state.hpp:19-122). */
enum { TITGPU_EOS_TAIT = 0, TITGPU_EOS_LINEAR_TAIT = 1 };
/* Time integrators (sph/time_integrator.hpp:32-228). */
enum { TITGPU_SYMPLECTIC_EULER = 0, TITGPU_VELOCITY_VERLET = 1, TITGPU_SSPRK2 = 2, TITGPU_SSPRK3 = 3 };

/* Replaces the construction of FluidEquations / integrator / ParticleArray
 * (wcsph.cpp:79-101): dimension (Space<Real,Dim>), kernel type, EOS type and
 * integrator type are the template arguments of the reference. */
TITGPU_API int titgpu_create(titgpu_ctx** ctx, int device, int dim, int kernel_id, int eos_id, int integrator_id);
TITGPU_API int titgpu_destroy(titgpu_ctx* ctx);
TITGPU_API const char* titgpu_last_error(const titgpu_ctx* ctx);

/* FluidEquations{g, mu, ..., TaitEquationOfState{cs0, rho0, xi}, Kernel{}} and
 * the uniform field h (fluid_equations.hpp:61-69, wcsph.cpp:116). The cell
 * hints of GridSearch{h}/GridFaceSearch{h} (wcsph.cpp:145-150) are accepted
 * for interface parity; the GPU hash always uses cells of one support radius
 * (neighbour sets do not depend on the cell size). */
TITGPU_API int titgpu_set_params(titgpu_ctx* ctx, double g, double mu, double cs0, double rho0, double xi, double h,
                      double search_cell_hint, double face_cell_hint);

/* The `Domain` (boundary-integral surface, normals into the fluid) and the
 * `Containment` surface (winding number +1 inside) of FluidEquations
 * (fluid_equations.hpp:528-529; wcsph.cpp:55-77). Vertex k of the domain
 * surface is fixed particle k (wcsph.cpp:110-113). Faces hold `dim` vertex
 * indices each. */
TITGPU_API int titgpu_set_surface(titgpu_ctx* ctx, const double* verts, size_t nv, const uint64_t* faces, size_t nf,
                       const double* inside_verts, size_t niv, const uint64_t* inside_faces, size_t nif);

/* ParticleArray::append + field assignment (particle_array.hpp:188-199,
 * 248-261). `field` is a reference field name: "r", "v", "rho", "m" are
 * state; "dv_dt" seeds the force time-step limit. A change of
 * (n_fluid, n_fixed) resets the particle set. */
TITGPU_API int titgpu_upload(titgpu_ctx* ctx, size_t n_fluid, size_t n_fixed, const char* field, const void* host, size_t stride_bytes);

/* field[particles] read-back (particle_array.hpp:264-276), any of the 17
 * varying fields of fluid_equations.hpp:41-48, original particle order. */
TITGPU_API int titgpu_download(titgpu_ctx* ctx, const char* field, void* host, size_t stride_bytes);

/* FluidEquations::initialize (fluid_equations.hpp:79-89). */
TITGPU_API int titgpu_initialize(titgpu_ctx* ctx);
/* FluidEquations::prepare (fluid_equations.hpp:99-105). */
TITGPU_API int titgpu_prepare(titgpu_ctx* ctx);
/* prepare + compute_continuity + compute_momentum (fluid_equations.hpp:232-305)
 * without advancing the state — for one-evaluation parity checks. */
TITGPU_API int titgpu_rhs_only(titgpu_ctx* ctx);
/* `nsteps` calls of <Integrator>::step(mesh, particles)
 * (time_integrator.hpp:49, 98, 161). Returns the last dt. The derived output
 * fields reflect the last step. */
TITGPU_API int titgpu_step(titgpu_ctx* ctx, int nsteps, double* dt_last);

/* Which derived fields titgpu_step publishes for titgpu_download after the LAST
 * step of a call. The reference recomputes all 17 varying fields of every
 * particle in every step (fluid_equations.hpp:331-512) although its time loop
 * reads them only when it writes an output frame (wcsph.cpp:186-189); on the
 * GPU publishing is optional:
 *   2 (default)  every field of every particle, as the reference;
 *   1            derived fields of fluid particles only (the shifting sums N, L,
 *                grad_v, grad_rho of wall particles are never read by the step);
 *   0            state only (r, v, rho, m); derived fields keep their last
 *                published values.
 * The state evolution is identical at every level. */
TITGPU_API int titgpu_set_outputs(titgpu_ctx* ctx, int level);

/* Step-persistent candidate lists (default OFF; titgpu_set_lists(ctx, 1) or the
 * environment TITGPU_LISTS=1 turns them on). The reference searches the neighbours at each of the four
 * prepare() calls of a step (sph/time_integrator.hpp:161-184). With lists the
 * cell sweep runs once per step with the radius enlarged by a skin of 0.1
 * radii, and every neighbour pass of the step applies the exact test to the
 * current positions of the listed candidates: the neighbour sets used are the
 * reference's. Should a particle move more than half the skin within a step
 * (beyond Mach 1/3 under the CFL limit) or a list overflow, the titgpu_step call
 * is repeated from its saved initial state with a search at every prepare;
 * titgpu_list_redos counts such calls. SSPRK integrators, single context.
 * Measured on B200 the lists do not pay (the pair passes are bound by the
 * gathers of the neighbour records, not by the cell sweep): kept as an option. */
TITGPU_API int titgpu_set_lists(titgpu_ctx* ctx, int on);
TITGPU_API unsigned long long titgpu_list_redos(const titgpu_ctx* ctx);

/* Whole steps as CUDA graphs (default ON; 2-D contexts without a slab decomposition, candidate
 * lists or profiling). The 2-D step has no host read-back, so its ~65 launches are recorded once
 * per state of the ping-pong buffers (they return to the same roles every third SSPRK step) and
 * replayed with one cudaGraphLaunch: the reference's default case (14 300 particles) is bound by
 * launch latency, not by any kernel. Results are bit-identical with graphs on or off.
 * TITGPU_GRAPHS=0 in the environment or titgpu_set_graphs(ctx, 0) turns them off. */
TITGPU_API int titgpu_set_graphs(titgpu_ctx* ctx, int on);
TITGPU_API unsigned long long titgpu_graph_replays(const titgpu_ctx* ctx);

/* Grouped candidate sweep of the kernel-sum passes (k_rhs_grp, k_shift_grp): one warp takes 4
 * consecutive particles of the cell order, sweeps their candidates once and keeps one
 * compacted hit list per particle; the pair sums then run per particle in the same order as
 * the default traversal, so the results are BIT-IDENTICAL. Faster on large particle counts
 * (k_rhs -13 % in 3-D, -22 % in 2-D at >= 1 M particles), slower on small ones (fewer, longer
 * warp tasks). mode: -1 = by particle count (default), 0 = never, 1 = always; environment
 * TITGPU_GROUP_SWEEP=auto|0|1. */
TITGPU_API int titgpu_set_group_sweep(titgpu_ctx* ctx, int mode);

/* Shared-memory-staged kernel-sum pass (3-D, kernels of support radius 2h; default OFF;
 * titgpu_set_tiles(ctx, 1) or the environment TITGPU_TILES=1 turns it on): one block per tile of
 * 2 x 2 x 2 search cells stages the 36 contiguous record runs of the 6 x 6 x 6 cells around it
 * in shared memory with cp.async.bulk + mbarrier and evaluates the pair sums from there
 * (csrc/tile.cuh). Same neighbour sets, same results to rounding (the order of the sums
 * differs). Measured on B200 it is slower than the default gather traversal (one block of 16
 * warps per SM cannot hide the FP64 latency): kept as an option. */
TITGPU_API int titgpu_set_tiles(titgpu_ctx* ctx, int on);

/* ParticleMesh adjacency (particle_mesh.hpp:67-72, 137-147): CSR, rows in
 * original particle order, columns ascending, self included. Call with
 * cols == NULL to obtain nnz. */
TITGPU_API int titgpu_neighbors(titgpu_ctx* ctx, uint64_t* row_offsets, uint64_t* cols, size_t cap, size_t* nnz);
/* ParticleMesh face adjacency, `mesh[domain, a]` (particle_mesh.hpp:74-82, 149-161;
 * geom/face_search/grid_face_search.hpp:91-112): per particle the indices of the domain
 * faces that intersect its support sphere, ascending. Same calling convention. */
TITGPU_API int titgpu_face_neighbors(titgpu_ctx* ctx, uint64_t* row_offsets, uint64_t* cols, size_t cap, size_t* nnz);

/* ---- Slab domain decomposition (one context per GPU / rank) -------------------
 * Replaces, across GPUs, what the reference's block partition does across
 * threads (sph/particle_mesh.hpp:165-241; geom/partition/sort_partition.hpp:25-75):
 * every rank owns the fluid particles of one slab [lo, hi) along `axis` plus GHOST
 * copies of the neighbouring slabs' particles within `halo` of its slab. Ghosts are
 * neighbours only: their right-hand sides are not evaluated. The rank-local wall
 * particles / faces are those of the slab plus halo and never move.
 *
 * Everything a step needs from the other ranks happens INSIDE titgpu_step, on the
 * context's stream (csrc/mg.cuh): migration of the particles that left the slab and
 * selection of the halo set at the first neighbour search of a step, a fixed-size
 * refresh of that set before each later search, {N, phi} and the shifted records inside
 * the shifting pass (fluid_equations.hpp:409-414, 443-450, 489-511), and a MIN / MAX
 * all-reduce of the time-step scalars (:203-221): ncclSend / ncclRecv / ncclAllReduce
 * between the GPUs of a box. A host in any language sets the slab, attaches a
 * communicator and calls titgpu_step; it never touches a ghost.
 *
 * Owned fluid records cross this part of the ABI as HOST arrays of 4 doubles per
 * particle in rank-local order:
 *   3-D: A = {x, y, z, rho}  B = {vx, vy, vz, m}
 *   2-D: A = {x, y, rho, m}  B = {vx, vy, 0, 0}
 * After the first step of a decomposed run the generic titgpu_upload / titgpu_download
 * address rank-local ids (owned, then ghosts, then walls), whose number changes from
 * step to step: use the _owned calls below instead. */

/* Capacity for the fluid particles of this rank (owned + ghosts); call before
 * the first titgpu_upload. */
TITGPU_API int titgpu_mg_reserve(titgpu_ctx* ctx, size_t max_fluid);
/* The slab of this rank: owned particles have lo <= r[axis] < hi (use -HUGE_VAL / HUGE_VAL
 * at the ends); `halo` = ghost-layer width including the margin for the motion within one
 * step (2 support radii + the longest wall-face edge + one particle spacing covers every
 * pass of the step). `fluid_total` >= 0 makes every step verify that the ranks together
 * still own that many fluid particles. An interior slab thinner than `halo` is refused. */
TITGPU_API int titgpu_mg_set_slab(titgpu_ctx* ctx, int axis, double lo, double hi, double halo, long long fluid_total);
/* Optional: away from the walls a neighbouring slab reads nothing beyond ONE support radius of its
 * own particles; the full `halo` is only needed for the fluid that can lie within a support radius
 * of a wall particle (the wall density is extrapolated from it, fluid_equations.hpp:122-164). With
 * `halo_pair` = support radius + margin (< halo) the ghost layer is that thin everywhere except along
 * the walls - about half the ghosts and half the bytes per exchange. 0 (default) = `halo` everywhere. */
TITGPU_API int titgpu_mg_set_halo_pair(titgpu_ctx* ctx, double halo_pair);
/* Global ids of the fluid particles uploaded so far (they travel with the particles). */
TITGPU_API int titgpu_mg_set_gids(titgpu_ctx* ctx, const int64_t* gids);
/* Communicator. Either adopt the host's ncclComm_t (rank r talks to r - 1 and r + 1), or
 * let the library create one from an ncclUniqueId (128 bytes) made on one rank by
 * titgpu_mg_nccl_unique_id and distributed by the host (MPI, torch.distributed, a file). */
TITGPU_API int titgpu_mg_attach_comm(titgpu_ctx* ctx, void* nccl_comm, int rank, int nranks);
TITGPU_API int titgpu_mg_nccl_unique_id(void* id128);
TITGPU_API int titgpu_mg_attach_nccl(titgpu_ctx* ctx, const void* id128, int rank, int nranks);
/* Ranks living in ONE process (one host thread per context, e.g. several ranks sharing a
 * GPU in tests): device-to-device copies ordered by CUDA events instead of NCCL, the same
 * exchange kernels. All ranks must call titgpu_step concurrently. */
TITGPU_API void* titgpu_mg_hub_create(int nranks);
TITGPU_API void titgpu_mg_hub_destroy(void* hub);
TITGPU_API int titgpu_mg_attach_hub(titgpu_ctx* ctx, void* hub, int rank);
TITGPU_API int titgpu_mg_detach(titgpu_ctx* ctx);
/* Current numbers of owned, ghost and wall particles. */
TITGPU_API int titgpu_mg_counts(titgpu_ctx* ctx, size_t* n_owned, size_t* n_ghost, size_t* n_fixed);
/* The owned particles (global ids and records, rank-local order) to / from HOST buffers;
 * any output pointer may be NULL. An upload drops the ghosts (the next step fetches them). */
TITGPU_API int titgpu_mg_download_owned(titgpu_ctx* ctx, int64_t* gid, double* A, double* B, size_t cap, size_t* n_owned);
TITGPU_API int titgpu_mg_upload_owned(titgpu_ctx* ctx, size_t n_owned, const int64_t* gid, const double* A, const double* B);
/* Exchanges performed and particles migrated away so far. */
TITGPU_API int titgpu_mg_stats(titgpu_ctx* ctx, unsigned long long* exchanges, unsigned long long* migrated);

/* Block until all queued work of the context has finished. */
TITGPU_API int titgpu_synchronize(titgpu_ctx* ctx);
/* Number of CUDA kernels this context has launched so far. */
TITGPU_API unsigned long long titgpu_launch_count(const titgpu_ctx* ctx);
/* The CUDA stream the context launches on (cudaStream_t), for event timing. */
TITGPU_API void* titgpu_stream(titgpu_ctx* ctx);
/* Per-kernel device timing (the analogue of the reference's TIT_PROFILE_SECTION
 * stopwatches, core/profiler.hpp:42-44): when enabled, every kernel launch is
 * bracketed by CUDA events on the context's stream and the durations are summed
 * per kernel name. `titgpu_profile_get` returns entry `index` (0 <=
 * index < titgpu_profile_count). */
TITGPU_API int titgpu_profile_enable(titgpu_ctx* ctx, int on);
TITGPU_API int titgpu_profile_reset(titgpu_ctx* ctx);
TITGPU_API int titgpu_profile_count(titgpu_ctx* ctx);
TITGPU_API int titgpu_profile_get(titgpu_ctx* ctx, int index, const char** name, unsigned long long* launches, double* total_ms);
/* Measured FP64 FMA throughput of the device (dependent-chain-free DFMA loop on
 * every SM), in TFLOP/s: the denominator of the FP64-pipe roofline. */
TITGPU_API int titgpu_measure_fp64_peak(titgpu_ctx* ctx, double* tflops);
/* Library version string. */
TITGPU_API const char* titgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TITGPU_H */
