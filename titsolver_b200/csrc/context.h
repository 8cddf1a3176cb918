// Host-side context of the B200 WCSPH engine (one context per GPU, one host
// thread per context). Owns every device buffer; the templated engine
// (engine.cuh) only launches kernels on them.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "common.cuh"

namespace titgpu {

// Field ids, in the reference's varying-field order
// (/root/reference/source/tit/sph/fluid_equations.hpp:41-48).
enum FieldId : int {
  F_m, F_gamma, F_grad_gamma, F_rho, F_drho_dt, F_grad_rho, F_p, F_cs, F_v, F_dv_dt, F_grad_v, F_r, F_dr, F_L, F_N, F_phi, F_rho_raw, F_COUNT
};
// 0 scalar, 1 vector, 2 matrix.
constexpr int kFieldRank[F_COUNT] = {0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 2, 1, 1, 2, 1, 0, 0};
constexpr const char* kFieldName[F_COUNT] = {"m", "gamma", "grad_gamma", "rho", "drho_dt", "grad_rho", "p", "cs", "v", "dv_dt", "grad_v", "r", "dr", "L", "N", "phi", "rho_raw"};

inline int field_width(int f, int dim) { return kFieldRank[f] == 0 ? 1 : kFieldRank[f] == 1 ? dim : dim * dim; }
inline int field_by_name(const char* s) {
  for (int f = 0; f < F_COUNT; ++f)
    if (std::string(s) == kFieldName[f]) return f;
  return -1;
}

struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t b) {
    if (b <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, b);
    if (e == cudaSuccess) bytes = b;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template<class T> T* as() const { return static_cast<T*>(p); }
};

struct EngineVTable;
struct MgTransport;

// Slab decomposition across GPUs (mg.cuh): the rank's slab, its transport and the
// buffers of the ghost-layer exchange.
struct MgState {
  MgTransport* tr = nullptr;  // owned; null = single context
  int axis = 0;
  double lo = -HUGE_VAL, hi = HUGE_VAL;  // owned slab [lo, hi) along `axis`
  double halo = 0;                       // ghost-layer width (incl. the margin for the motion within a step)
  double halo_pair = 0;                  // ... of which away from the walls only this much is needed (0 = all of it)
  int left = -1, right = -1;             // neighbour ranks (-1 = none)
  // The halo set of the current step: n_send[s] owned particles go to side s (0 left, 1 right),
  // n_recv[s] ghosts come from it. Buffers hold [left part | right part].
  size_t n_send[2] = {0, 0}, n_recv[2] = {0, 0};
  bool set_valid = false;
  DBuf sendA, sendB, recvA, recvB, send_idx;
  DBuf migA[2], migB[2], migG[2];  // records / global ids of the particles leaving to side s
  DBuf gid, gid_alt;               // global id by local id (owned particles)
  DBuf flags, scans, pos_of, cub_tmp, bad;
  DBuf stageA, stageB;             // owned records in local-id order (titgpu_mg_download_owned / _upload_owned)
  unsigned long long exchanges = 0, migrated = 0;
  long long fluid_total = -1;  // global number of fluid particles (conservation check), -1 = unchecked
};

struct Ctx {
  int device = 0, dim = 2, kernel_id = 4, eos_id = 0, integrator_id = 3;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  const EngineVTable* vt = nullptr;
  Params prm{};
  bool params_set = false, surface_set = false, sized = false, grid_ready = false, initialized = false;
  bool sorted_identity = true;  // state order == original order
  size_t nf = 0, nx = 0, n = 0;
  double search_hint = 0, face_hint = 0;
  // Slab decomposition (titgpu_mg_*): capacity reserved for a varying number of
  // fluid particles, and the callback the step invokes where ranks must talk.
  size_t reserve_fluid = 0, cap_n = 0;
  MgState mg;
  // Step-persistent candidate lists (engine.cuh, "candidate lists"): built once
  // per step with a skin, reused by every neighbour pass of the step.
  bool lists_enabled = false;  // titgpu_set_lists / TITGPU_LISTS=1 (measured: no gain on B200)
  bool lists_active = false;   // valid for the current particle order
  bool force_safe = false;     // redo in progress: search at every prepare, as the reference does
  int nl_stride = 0;
  DBuf nl_idx, nl_cnt;
  DBuf bakA, bakB, bak_orig;   // state at the beginning of a titgpu_step call (redo)
  unsigned long long list_redos = 0;  // titgpu_step calls that had to be repeated without lists
  int output_level = 2;  // titgpu_set_outputs: 0 state only, 1 + derived fields of fluid particles, 2 all (reference)

  // Packed particle records in sorted order (see engine.cuh): A = position +
  // density (+ mass in 2-D), B = velocity (+ mass in 3-D); A0 / B0 = the state
  // at the beginning of the step (SSPRK). Each exists twice (ping-pong for the
  // cell reorder and for the fused RHS + update pass); the pointers swap
  // independently.
  DBuf bufA[4], bufB[4], buf_orig[2];
  double4 *A = nullptr, *B = nullptr, *A_alt = nullptr, *B_alt = nullptr;
  double4 *A0 = nullptr, *B0 = nullptr, *A0_alt = nullptr, *B0_alt = nullptr;
  int *orig = nullptr, *orig_alt = nullptr;

  // Per-sorted-particle derived data.
  DBuf C;                        // {cs, p / rho^2, 1 / rho, p}
  DBuf F;                        // float4 grid coordinates + flags (FP32 pre-filter)
  DBuf gamma_w, gg_w, wsum;      // wall pass: gamma, grad gamma, face sums of the consumer
  DBuf gamma_s, N_s, phi_s, phi2_s, dr_s, gv_s, gr_s, fs_flag;  // post-integration scratch

  // Tiles of the shared-memory-staged pair passes (tile.cuh): non-empty tiles of the current sort.
  bool tiles_enabled = false, tiles_valid = false;  // TITGPU_TILES=1 / titgpu_set_tiles: measured slower than the gather traversal on B200, kept as an option
  DBuf tile_list, tile_count;                      // tile ids | {count, work cursor}
  int tile_nty = 0, tile_ntz = 0;

  // CUDA graphs of whole steps (2-D: the step has no host read-back). One graph per state of
  // the ping-pong buffers (they return to the same roles every third SSPRK step) and output
  // flag; dropped whenever the particle set, the grid or the parameters change.
  struct StepGraph {
    void* key[10];   // A, A_alt, A0, A0_alt, B, B_alt, B0, B0_alt, orig, orig_alt before the step
    void* after[10]; // ... and after it
    int write_out, output_level;
    unsigned long long launches;
    cudaGraphExec_t exec;
  };
  int group_sweep = -1;        // grouped candidate sweep (k_rhs_grp / k_shift_grp): -1 = by size (n >= kGroupMinN), 0 = never, 1 = always (titgpu_set_group_sweep / TITGPU_GROUP_SWEEP)
  bool graphs_enabled = true;  // TITGPU_GRAPHS=0 / titgpu_set_graphs
  std::vector<StepGraph> graphs;
  unsigned long long graph_replays = 0;
  void drop_graphs() {
    for (StepGraph& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  // Hash / sort scratch.
  DBuf cell_id, slot, tmp_perm, perm, cell_cnt, cell_start, cub_tmp, cell_fs, cell_fluid;

  // Static boundary.
  DBuf frames, fcell_start, fcell_faces, face_cells, fflag, ftwin, fterm, fgeom, favg;
  // 3-D wall pipeline work lists (engine.cuh, k_wsearch) and their capacities in entries.
  DBuf ww_faces, ww_sref, ww_items, ww_val, ww_rims, ww_val2, ww_act, ww_ovf, ww_x2, ww_cur, ww_list;
  size_t ww_cap_faces = 0, ww_cap_items = 0, ww_cap_rims = 0, ww_cap_act = 0;
  size_t nfaces = 0;
  DBuf cverts, cfaces;
  size_t ncfaces = 0;
  DBuf gamma_fixed, gg_fixed;  // static gamma / grad gamma of the fixed particles
  DBuf rho_fx, p_fx;           // wall density / pressure by fixed id
  bool fixed_cache_valid = false;
  // Published fields of wall particles far from any fluid are static (engine.cuh, k_deep_dry).
  double wall_edge_max = 0;  // longest bounding-box diagonal of a wall face (setup_grid)
  DBuf dry_pub, dry_skip;
  bool dry_pub_valid = false, dry_cache_enabled = true;  // TITGPU_DRY_CACHE=0 recomputes them at every publishing step

  // Outputs, original particle order, one buffer per field.
  DBuf out[F_COUNT];
  DBuf staging;

  // Device scalars: [0] dt, [1] max |dv_dt|^2 of the last RHS (bits), [2] dt
  // reduction (bits), [3] spare, [4] candidate-list flags, [5] backup of [1],
  // [6] (int) set by an upload of `r` that moves a wall particle.
  DBuf scalars;

  // Host copies of the surfaces.
  std::vector<double> h_verts, h_cverts;
  std::vector<uint64_t> h_faces, h_cfaces;

  // Counters.
  unsigned long long launches = 0;

  // Optional per-kernel timing (titgpu_profile_*): CUDA events on `stream`
  // around every launch, folded into per-name totals after each API call.
  struct ProfPending { const char* name; cudaEvent_t beg, end; };
  struct ProfTotal { std::string name; unsigned long long count = 0; double ms = 0; };
  bool prof_on = false;
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<ProfTotal> prof_totals;
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  cudaEvent_t prof_begin(const char* name) {
    if (!prof_on) return nullptr;
    ProfPending p{name, prof_event(), prof_event()};
    cudaEventRecord(p.beg, stream);
    prof_pending.push_back(p);
    return p.end;
  }
  void prof_end(cudaEvent_t e) { if (e) cudaEventRecord(e, stream); }
  // Requires the stream to be idle (call after cudaStreamSynchronize).
  void prof_fold() {
    for (const ProfPending& p : prof_pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.beg, p.end) == cudaSuccess) {
        ProfTotal* t = nullptr;
        for (ProfTotal& q : prof_totals) if (q.name == p.name) { t = &q; break; }
        if (!t) { prof_totals.push_back(ProfTotal{p.name}); t = &prof_totals.back(); }
        t->count++;
        t->ms += ms;
      }
      prof_pool.push_back(p.beg);
      prof_pool.push_back(p.end);
    }
    prof_pending.clear();
  }

  std::string err;
};

// Per-(dim, kernel) entry points, filled by the explicit instantiations.
struct EngineVTable {
  void (*fill_params)(Ctx&);
  int (*seed_fmax)(Ctx&);
  int (*set_surface)(Ctx&);
  int (*initialize)(Ctx&);
  int (*prepare)(Ctx&, bool write_out);
  int (*rhs_only)(Ctx&);
  int (*step)(Ctx&, int nsteps);
  int (*neighbors)(Ctx&, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz);
  int (*face_neighbors)(Ctx&, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz);
  int (*download_state)(Ctx&, int field, double* dst_dev);  // unsort into original order
  int (*upload_state)(Ctx&, int field, const double* src_dev);
  // Slab decomposition: owned records in local-id order to / from device staging buffers.
  int (*mg_gather_owned)(Ctx&, double* A_dev, double* B_dev);
  int (*mg_replace_owned)(Ctx&, size_t n_owned, const double* A_dev, const double* B_dev);
};

const EngineVTable* get_engine(int dim, int kernel_id);
void register_engine(int dim, int kernel_id, const EngineVTable* vt);

#define TIT_CUDA_OK(ctx, expr)                                                                     \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      (ctx).err = std::string(#expr) + ": " + cudaGetErrorString(e_);                              \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

}  // namespace titgpu
