// extern "C" boundary of libtitgpu.so (see include/titgpu.h). Plain pointers
// and sizes only; dispatches to the (dimension, kernel) engine instantiation.
#include "../../include/titgpu.h"

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "context.h"
#include "mg_transport.h"

struct titgpu_ctx {
  titgpu::Ctx c;
};

namespace titgpu {

namespace {
const EngineVTable* g_engines[4][6] = {};
}
void register_engine(int dim, int kernel_id, const EngineVTable* vt) {
  if (dim >= 2 && dim <= 3 && kernel_id >= 0 && kernel_id < 6) g_engines[dim][kernel_id] = vt;
}
const EngineVTable* get_engine(int dim, int kernel_id) {
  if (dim < 2 || dim > 3 || kernel_id < 0 || kernel_id >= 6) return nullptr;
  return g_engines[dim][kernel_id];
}

namespace {

int fail(Ctx& c, const std::string& msg) {
  c.err = msg;
  return 1;
}

int alloc_particles(Ctx& c) {
  // Buffers are sized for the reserved capacity (slab decomposition: the number
  // of fluid particles of a rank varies from exchange to exchange).
  c.cap_n = std::max<size_t>(std::max(c.nf, c.reserve_fluid) + c.nx, 1);
  const size_t n = c.cap_n, D = size_t(c.dim);
  for (DBuf& b : c.bufA) TIT_CUDA_OK(c, b.ensure(n * sizeof(double4)));
  for (DBuf& b : c.bufB) TIT_CUDA_OK(c, b.ensure(n * sizeof(double4)));
  for (DBuf& b : c.buf_orig) TIT_CUDA_OK(c, b.ensure(n * 4));
  c.A = c.bufA[0].as<double4>(); c.A_alt = c.bufA[1].as<double4>(); c.A0 = c.bufA[2].as<double4>(); c.A0_alt = c.bufA[3].as<double4>();
  c.B = c.bufB[0].as<double4>(); c.B_alt = c.bufB[1].as<double4>(); c.B0 = c.bufB[2].as<double4>(); c.B0_alt = c.bufB[3].as<double4>();
  c.orig = c.buf_orig[0].as<int>(); c.orig_alt = c.buf_orig[1].as<int>();
  TIT_CUDA_OK(c, c.C.ensure(n * sizeof(double4)));
  TIT_CUDA_OK(c, c.F.ensure(n * sizeof(float4)));
  for (DBuf* b : {&c.gamma_w, &c.gamma_s, &c.phi_s, &c.phi2_s}) TIT_CUDA_OK(c, b->ensure(n * 8));
  for (DBuf* b : {&c.gg_w, &c.N_s, &c.dr_s, &c.gr_s}) TIT_CUDA_OK(c, b->ensure(n * D * 8));
  TIT_CUDA_OK(c, c.gv_s.ensure(n * D * D * 8));
  TIT_CUDA_OK(c, c.wsum.ensure(n * (2 * D + 2 * D * D) * 8));
  TIT_CUDA_OK(c, c.fs_flag.ensure(n));
  for (DBuf* b : {&c.cell_id, &c.slot, &c.tmp_perm, &c.perm}) TIT_CUDA_OK(c, b->ensure(n * 4));
  const size_t nx = std::max<size_t>(c.nx, 1);
  TIT_CUDA_OK(c, c.gamma_fixed.ensure(nx * 8));
  TIT_CUDA_OK(c, c.gg_fixed.ensure(nx * D * 8));
  TIT_CUDA_OK(c, c.rho_fx.ensure(nx * 8));
  TIT_CUDA_OK(c, c.p_fx.ensure(nx * 8));
  for (int f = 0; f < F_COUNT; ++f) {
    const size_t bytes = n * size_t(field_width(f, c.dim)) * 8;
    TIT_CUDA_OK(c, c.out[f].ensure(bytes));
    TIT_CUDA_OK(c, cudaMemsetAsync(c.out[f].p, 0, bytes, c.stream));
  }
  TIT_CUDA_OK(c, c.staging.ensure(n * D * D * 8));
  TIT_CUDA_OK(c, c.scalars.ensure(64));
  TIT_CUDA_OK(c, cudaMemsetAsync(c.scalars.p, 0, 64, c.stream));
  // Zero state, identity order.
  for (DBuf& b : c.bufA) TIT_CUDA_OK(c, cudaMemsetAsync(b.p, 0, n * sizeof(double4), c.stream));
  for (DBuf& b : c.bufB) TIT_CUDA_OK(c, cudaMemsetAsync(b.p, 0, n * sizeof(double4), c.stream));
  std::vector<int> id(c.n);
  for (size_t i = 0; i < c.n; ++i) id[i] = int(i);
  if (c.n) TIT_CUDA_OK(c, cudaMemcpyAsync(c.orig, id.data(), c.n * 4, cudaMemcpyHostToDevice, c.stream));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  c.sorted_identity = true;
  c.grid_ready = false;
  c.fixed_cache_valid = false;
  c.drop_graphs();
  c.dry_pub_valid = false;
  c.sized = true;
  c.prm.nf = int(c.nf); c.prm.nx = int(c.nx); c.prm.n = int(c.n); c.prm.n_owned = int(c.nf);
  return 0;
}

// Host <-> packed conversion honouring stride_bytes.
void pack_in(const void* host, size_t n, int width, size_t stride, std::vector<double>& out) {
  out.resize(n * width);
  const char* p = static_cast<const char*>(host);
  for (size_t i = 0; i < n; ++i) std::memcpy(&out[i * width], p + i * stride, size_t(width) * 8);
}
void pack_out(const std::vector<double>& in, size_t n, int width, size_t stride, void* host) {
  char* p = static_cast<char*>(host);
  for (size_t i = 0; i < n; ++i) std::memcpy(p + i * stride, &in[i * width], size_t(width) * 8);
}

int check_ready(Ctx& c, bool need_particles) {
  if (!c.params_set) return fail(c, "titgpu_set_params has not been called");
  if (need_particles && !c.sized) return fail(c, "no particles uploaded");
  if (need_particles && c.cap_n > 0x7fffffffull / 16) return fail(c, "too many particles for 32-bit indexing");
  return 0;
}

}  // namespace
}  // namespace titgpu

using namespace titgpu;

extern "C" {

const char* titgpu_version(void) { return "titgpu 0.1 (sm_100a)"; }

int titgpu_create(titgpu_ctx** out, int device, int dim, int kernel_id, int eos_id, int integrator_id) {
  if (!out) return 1;
  *out = nullptr;
  titgpu_ctx* h = new (std::nothrow) titgpu_ctx();
  if (!h) return 1;
  *out = h;  // returned even on failure so that titgpu_last_error() works
  Ctx& c = h->c;
  c.device = device; c.dim = dim; c.kernel_id = kernel_id; c.eos_id = eos_id; c.integrator_id = integrator_id;
  c.vt = get_engine(dim, kernel_id);
  if (!c.vt) return fail(c, "no engine built for this (dim, kernel_id)");
  if (eos_id < 0 || eos_id > 1) return fail(c, "bad eos_id");
  if (integrator_id < 0 || integrator_id > 3) return fail(c, "bad integrator_id");
  TIT_CUDA_OK(c, cudaSetDevice(device));
  TIT_CUDA_OK(c, cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  TIT_CUDA_OK(c, cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
  c.prm.eos = eos_id;
  if (const char* e = std::getenv("TITGPU_LISTS")) c.lists_enabled = e[0] != '0';
  if (const char* e = std::getenv("TITGPU_TILES")) c.tiles_enabled = e[0] != '0';
  if (const char* e = std::getenv("TITGPU_GRAPHS")) c.graphs_enabled = e[0] != '0';
  if (const char* e = std::getenv("TITGPU_DRY_CACHE")) c.dry_cache_enabled = e[0] != '0';
  if (const char* e = std::getenv("TITGPU_GROUP_SWEEP")) c.group_sweep = e[0] == 'a' ? -1 : e[0] != '0';
  return 0;
}

int titgpu_destroy(titgpu_ctx* h) {
  if (!h) return 0;
  Ctx& c = h->c;
  if (c.stream) { cudaSetDevice(c.device); cudaStreamSynchronize(c.stream); }
  for (DBuf& b : c.bufA) b.release();
  for (DBuf& b : c.bufB) b.release();
  for (DBuf& b : c.buf_orig) b.release();
  for (DBuf* b : {&c.C, &c.F, &c.gamma_w, &c.gg_w, &c.wsum, &c.gamma_s, &c.N_s, &c.phi_s, &c.phi2_s, &c.dr_s, &c.gv_s, &c.gr_s, &c.fs_flag, &c.cell_id, &c.slot, &c.tmp_perm, &c.perm,
                  &c.cell_cnt, &c.cell_start, &c.cub_tmp, &c.cell_fs, &c.cell_fluid, &c.frames, &c.fcell_start, &c.fcell_faces, &c.face_cells, &c.fflag, &c.ftwin, &c.fterm, &c.fgeom, &c.favg, &c.ww_faces, &c.ww_sref, &c.ww_items, &c.ww_val, &c.ww_rims, &c.ww_val2, &c.ww_act, &c.ww_ovf, &c.ww_x2, &c.ww_cur, &c.ww_list, &c.cverts, &c.cfaces, &c.gamma_fixed, &c.gg_fixed,
                  &c.rho_fx, &c.p_fx, &c.dry_pub, &c.dry_skip, &c.staging, &c.scalars, &c.tile_list, &c.tile_count, &c.nl_idx, &c.nl_cnt, &c.bakA, &c.bakB, &c.bak_orig})
    b->release();
  for (cudaEvent_t e : c.prof_pool) cudaEventDestroy(e);
  for (auto& p : c.prof_pending) { cudaEventDestroy(p.beg); cudaEventDestroy(p.end); }
  for (DBuf& b : c.out) b.release();
  c.drop_graphs();
  delete c.mg.tr;
  c.mg.tr = nullptr;
  for (DBuf* b : {&c.mg.sendA, &c.mg.sendB, &c.mg.recvA, &c.mg.recvB, &c.mg.send_idx, &c.mg.migA[0], &c.mg.migA[1], &c.mg.migB[0], &c.mg.migB[1], &c.mg.migG[0], &c.mg.migG[1], &c.mg.gid,
                  &c.mg.gid_alt, &c.mg.flags, &c.mg.scans, &c.mg.pos_of, &c.mg.cub_tmp, &c.mg.bad, &c.mg.stageA, &c.mg.stageB})
    b->release();
  if (c.stream) cudaStreamDestroy(c.stream);
  delete h;
  return 0;
}

const char* titgpu_last_error(const titgpu_ctx* h) { return h ? h->c.err.c_str() : "null context"; }

int titgpu_set_params(titgpu_ctx* h, double g, double mu, double cs0, double rho0, double xi, double hh, double search_hint, double face_hint) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (!c.vt) return fail(c, "context has no engine");
  if (!(hh > 0)) return fail(c, "kernel width h must be positive");
  if (!(cs0 > 0) || !(rho0 > 0)) return fail(c, "cs0 and rho0 must be positive");
  Params& P = c.prm;
  P.g = g; P.mu = mu; P.cs0 = cs0; P.rho0 = rho0; P.xi = xi; P.h = hh;
  P.eos = c.eos_id;
  c.search_hint = search_hint; c.face_hint = face_hint;
  c.vt->fill_params(c);
  c.params_set = true;
  c.grid_ready = false;
  c.fixed_cache_valid = false;
  c.dry_pub_valid = false;
  return 0;
}

int titgpu_set_surface(titgpu_ctx* h, const double* verts, size_t nv, const uint64_t* faces, size_t nf, const double* iv, size_t niv, const uint64_t* ifc, size_t nif) {
  if (!h) return 1;
  Ctx& c = h->c;
  const size_t D = size_t(c.dim);
  for (size_t i = 0; i < nf * D; ++i)
    if (faces[i] >= nv) return fail(c, "surface face refers to a vertex out of range");
  for (size_t i = 0; i < nif * D; ++i)
    if (ifc[i] >= niv) return fail(c, "containment face refers to a vertex out of range");
  c.h_verts.assign(verts, verts + nv * D);
  c.h_faces.assign(faces, faces + nf * D);
  c.h_cverts.assign(iv, iv + niv * D);
  c.h_cfaces.assign(ifc, ifc + nif * D);
  c.surface_set = true;
  c.grid_ready = false;
  c.fixed_cache_valid = false;
  c.dry_pub_valid = false;
  return 0;
}

int titgpu_upload(titgpu_ctx* h, size_t n_fluid, size_t n_fixed, const char* field, const void* host, size_t stride) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (!c.vt) return fail(c, "context has no engine");
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  const int f = field_by_name(field);
  if (f < 0) return fail(c, std::string("unknown field '") + field + "'");
  if (!c.sized || n_fluid != c.nf || n_fixed != c.nx) {
    c.nf = n_fluid; c.nx = n_fixed; c.n = n_fluid + n_fixed;
    if (std::max(c.nf, c.reserve_fluid) + c.nx > 0x7fffffffull / 16) return fail(c, "too many particles for 32-bit indexing");
    if (alloc_particles(c)) return 1;
  }
  if (c.n == 0) return 0;
  const int w = field_width(f, c.dim);
  if (stride == 0) stride = size_t(w) * 8;
  if (stride < size_t(w) * 8) return fail(c, "stride_bytes smaller than the field value");
  std::vector<double> packed;
  const double* src = static_cast<const double*>(host);
  if (stride != size_t(w) * 8) { pack_in(host, c.n, w, stride, packed); src = packed.data(); }
  const size_t bytes = c.n * size_t(w) * 8;
  const bool is_state = (f == F_r || f == F_v || f == F_rho || f == F_m);
  // an upload may change what the wall particles' published fields were computed from (or the fields themselves)
  if (!(f == F_r && c.fixed_cache_valid) && f != F_dv_dt) c.dry_pub_valid = false;
  if (!is_state) {
    TIT_CUDA_OK(c, cudaMemcpyAsync(c.out[f].p, src, bytes, cudaMemcpyHostToDevice, c.stream));
    if (f == F_dv_dt && c.vt->seed_fmax(c)) return 1;
    TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return 0;
  }
  TIT_CUDA_OK(c, cudaMemcpyAsync(c.staging.p, src, bytes, cudaMemcpyHostToDevice, c.stream));
  // Wall particles never move in the reference's time loop (wcsph.cpp:170-193): their
  // gamma / grad gamma cache survives an upload of `r` that leaves them where they were.
  const bool track_walls = f == F_r && c.fixed_cache_valid && c.grid_ready;
  if (track_walls) TIT_CUDA_OK(c, cudaMemsetAsync(c.scalars.as<int>() + 12, 0, 4, c.stream));
  if (c.vt->upload_state(c, f, c.staging.as<double>())) return 1;
  int moved = 1;
  if (track_walls) TIT_CUDA_OK(c, cudaMemcpyAsync(&moved, c.scalars.as<int>() + 12, 4, cudaMemcpyDeviceToHost, c.stream));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  c.prof_fold();
  if (f == F_r && moved) { c.fixed_cache_valid = false; c.dry_pub_valid = false; }
  return 0;
}

int titgpu_download(titgpu_ctx* h, const char* field, void* host, size_t stride) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (!c.sized) return fail(c, "no particles uploaded");
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  const int f = field_by_name(field);
  if (f < 0) return fail(c, std::string("unknown field '") + field + "'");
  if (c.n == 0) return 0;
  const int w = field_width(f, c.dim);
  if (stride == 0) stride = size_t(w) * 8;
  if (stride < size_t(w) * 8) return fail(c, "stride_bytes smaller than the field value");
  const size_t bytes = c.n * size_t(w) * 8;
  const bool is_state = (f == F_r || f == F_v || f == F_rho || f == F_m);
  const double* src = c.out[f].as<double>();
  if (is_state) {
    if (c.vt->download_state(c, f, c.staging.as<double>())) return 1;
    src = c.staging.as<double>();
  }
  if (stride == size_t(w) * 8) {
    TIT_CUDA_OK(c, cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  } else {
    std::vector<double> packed(c.n * w);
    TIT_CUDA_OK(c, cudaMemcpyAsync(packed.data(), src, bytes, cudaMemcpyDeviceToHost, c.stream));
    TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    pack_out(packed, c.n, w, stride, host);
  }
  return 0;
}

#define TITGPU_ENTER(need_particles)                 \
  if (!h) return 1;                                  \
  Ctx& c = h->c;                                     \
  if (!c.vt) return fail(c, "context has no engine"); \
  if (check_ready(c, need_particles)) return 1;      \
  TIT_CUDA_OK(c, cudaSetDevice(c.device));

#define TITGPU_LEAVE()                                   \
  TIT_CUDA_OK(c, cudaGetLastError());                    \
  c.prof_fold();                                         \
  return 0;

int titgpu_initialize(titgpu_ctx* h) {
  TITGPU_ENTER(true)
  if (c.vt->initialize(c)) return 1;
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  TITGPU_LEAVE()
}
int titgpu_prepare(titgpu_ctx* h) {
  TITGPU_ENTER(true)
  if (c.vt->prepare(c, true)) return 1;
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  TITGPU_LEAVE()
}
int titgpu_rhs_only(titgpu_ctx* h) {
  TITGPU_ENTER(true)
  if (c.vt->rhs_only(c)) return 1;
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  TITGPU_LEAVE()
}
int titgpu_step(titgpu_ctx* h, int nsteps, double* dt_last) {
  TITGPU_ENTER(true)
  if (nsteps < 0) return fail(c, "nsteps must be non-negative");
  if (c.vt->step(c, nsteps)) return 1;
  double dt = 0;
  TIT_CUDA_OK(c, cudaMemcpyAsync(&dt, c.scalars.p, 8, cudaMemcpyDeviceToHost, c.stream));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  if (dt_last) *dt_last = dt;
  TITGPU_LEAVE()
}
int titgpu_set_lists(titgpu_ctx* h, int on) {
  if (!h) return 1;
  Ctx& c = h->c;
  c.lists_enabled = on != 0;
  c.lists_active = false;
  c.grid_ready = false;  // the cell size depends on the skin
  return 0;
}
unsigned long long titgpu_list_redos(const titgpu_ctx* h) { return h ? h->c.list_redos : 0; }
int titgpu_set_graphs(titgpu_ctx* h, int on) {
  if (!h) return 1;
  h->c.graphs_enabled = on != 0;
  if (!on) h->c.drop_graphs();
  return 0;
}
unsigned long long titgpu_graph_replays(const titgpu_ctx* h) { return h ? h->c.graph_replays : 0; }
int titgpu_set_group_sweep(titgpu_ctx* h, int mode) {
  if (!h) return 1;
  h->c.group_sweep = mode < 0 ? -1 : mode != 0;
  return 0;
}
int titgpu_set_tiles(titgpu_ctx* h, int on) {
  if (!h) return 1;
  h->c.tiles_enabled = on != 0;
  h->c.tiles_valid = false;
  return 0;
}
int titgpu_set_outputs(titgpu_ctx* h, int level) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (level < 0 || level > 2) return fail(c, "output level must be 0, 1 or 2");
  c.output_level = level;
  return 0;
}
// ---- slab decomposition --------------------------------------------------------
int titgpu_mg_reserve(titgpu_ctx* h, size_t max_fluid) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (c.sized && max_fluid + c.nx > c.cap_n) return fail(c, "titgpu_mg_reserve must precede the first upload");
  c.reserve_fluid = max_fluid;
  return 0;
}
int titgpu_mg_counts(titgpu_ctx* h, size_t* n_owned, size_t* n_ghost, size_t* n_fixed) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (n_owned) *n_owned = size_t(c.prm.n_owned);
  if (n_ghost) *n_ghost = c.nf - size_t(c.prm.n_owned);
  if (n_fixed) *n_fixed = c.nx;
  return 0;
}
int titgpu_mg_set_halo_pair(titgpu_ctx* h, double halo_pair) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (!(halo_pair >= 0)) return fail(c, "titgpu_mg_set_halo_pair: negative width");
  c.mg.halo_pair = halo_pair;
  c.mg.set_valid = false;
  return 0;
}
int titgpu_mg_set_slab(titgpu_ctx* h, int axis, double lo, double hi, double halo, long long fluid_total) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (axis < 0 || axis >= c.dim) return fail(c, "titgpu_mg_set_slab: bad axis");
  if (!(lo < hi)) return fail(c, "titgpu_mg_set_slab: empty slab");
  if (!(halo > 0)) return fail(c, "titgpu_mg_set_slab: the halo width must be positive");
  // Ghosts come from the adjacent slabs only: an interior slab thinner than the halo would
  // need particles of the slab after next (and could lose particles that cross it in one step).
  if (std::isfinite(lo) && std::isfinite(hi) && hi - lo < halo) return fail(c, "titgpu_mg_set_slab: interior slab thinner than the halo (use fewer ranks)");
  c.mg.axis = axis; c.mg.lo = lo; c.mg.hi = hi; c.mg.halo = halo; c.mg.fluid_total = fluid_total;
  c.mg.set_valid = false;
  return 0;
}
int titgpu_mg_set_gids(titgpu_ctx* h, const int64_t* gids) {
  TITGPU_ENTER(true)
  if (!gids) return fail(c, "titgpu_mg_set_gids: null");
  TIT_CUDA_OK(c, c.mg.gid.ensure(c.cap_n * 8));
  if (c.nf) TIT_CUDA_OK(c, cudaMemcpyAsync(c.mg.gid.p, gids, c.nf * 8, cudaMemcpyHostToDevice, c.stream));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  return 0;
}
namespace {
int mg_attach(Ctx& c, MgTransport* t, const std::string& err) {
  if (!t) return fail(c, err);
  delete c.mg.tr;
  c.mg.tr = t;
  c.mg.left = t->rank() > 0 ? t->rank() - 1 : -1;
  c.mg.right = t->rank() + 1 < t->nranks() ? t->rank() + 1 : -1;
  c.mg.set_valid = false;
  return 0;
}
}  // namespace
int titgpu_mg_nccl_unique_id(void* id128) {
  std::string err;
  return id128 ? nccl_unique_id(id128, err) : 1;
}
int titgpu_mg_attach_nccl(titgpu_ctx* h, const void* id128, int rank, int nranks) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (!id128 || rank < 0 || rank >= nranks) return fail(c, "titgpu_mg_attach_nccl: bad arguments");
  std::string err;
  return mg_attach(c, make_nccl_transport(id128, rank, nranks, c.device, err), err);
}
int titgpu_mg_attach_comm(titgpu_ctx* h, void* nccl_comm, int rank, int nranks) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (rank < 0 || rank >= nranks) return fail(c, "titgpu_mg_attach_comm: bad arguments");
  std::string err;
  return mg_attach(c, adopt_nccl_comm(nccl_comm, rank, nranks, err), err);
}
void* titgpu_mg_hub_create(int nranks) { return make_hub(nranks); }
void titgpu_mg_hub_destroy(void* hub) { destroy_hub(static_cast<MgHub*>(hub)); }
int titgpu_mg_attach_hub(titgpu_ctx* h, void* hub, int rank) {
  if (!h) return 1;
  Ctx& c = h->c;
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  std::string err;
  return mg_attach(c, make_hub_transport(static_cast<MgHub*>(hub), rank, err), err);
}
int titgpu_mg_detach(titgpu_ctx* h) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (c.stream) { cudaSetDevice(c.device); cudaStreamSynchronize(c.stream); }
  delete c.mg.tr;
  c.mg.tr = nullptr;
  c.mg.left = c.mg.right = -1;
  return 0;
}
int titgpu_mg_download_owned(titgpu_ctx* h, int64_t* gid, double* A, double* B, size_t cap, size_t* n_owned) {
  TITGPU_ENTER(true)
  const size_t no = size_t(c.prm.n_owned);
  if (n_owned) *n_owned = no;
  if (!A && !B && !gid) return 0;
  if (cap < no) return fail(c, "titgpu_mg_download_owned: capacity too small");
  TIT_CUDA_OK(c, c.mg.stageA.ensure(std::max<size_t>(c.cap_n, 1) * sizeof(double4)));
  TIT_CUDA_OK(c, c.mg.stageB.ensure(std::max<size_t>(c.cap_n, 1) * sizeof(double4)));
  if (c.vt->mg_gather_owned(c, c.mg.stageA.as<double>(), c.mg.stageB.as<double>())) return 1;
  if (no) {
    if (A) TIT_CUDA_OK(c, cudaMemcpyAsync(A, c.mg.stageA.p, no * sizeof(double4), cudaMemcpyDeviceToHost, c.stream));
    if (B) TIT_CUDA_OK(c, cudaMemcpyAsync(B, c.mg.stageB.p, no * sizeof(double4), cudaMemcpyDeviceToHost, c.stream));
    if (gid) {
      if (c.mg.gid.bytes >= no * 8) TIT_CUDA_OK(c, cudaMemcpyAsync(gid, c.mg.gid.p, no * 8, cudaMemcpyDeviceToHost, c.stream));
      else for (size_t i = 0; i < no; ++i) gid[i] = int64_t(i);
    }
  }
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  TITGPU_LEAVE()
}
int titgpu_mg_upload_owned(titgpu_ctx* h, size_t n_owned, const int64_t* gid, const double* A, const double* B) {
  TITGPU_ENTER(true)
  if (n_owned && (!A || !B)) return fail(c, "titgpu_mg_upload_owned: null records");
  if (n_owned + c.nx > c.cap_n) return fail(c, "titgpu_mg_upload_owned: more particles than reserved (titgpu_mg_reserve)");
  TIT_CUDA_OK(c, c.mg.stageA.ensure(std::max<size_t>(c.cap_n, 1) * sizeof(double4)));
  TIT_CUDA_OK(c, c.mg.stageB.ensure(std::max<size_t>(c.cap_n, 1) * sizeof(double4)));
  if (n_owned) {
    TIT_CUDA_OK(c, cudaMemcpyAsync(c.mg.stageA.p, A, n_owned * sizeof(double4), cudaMemcpyHostToDevice, c.stream));
    TIT_CUDA_OK(c, cudaMemcpyAsync(c.mg.stageB.p, B, n_owned * sizeof(double4), cudaMemcpyHostToDevice, c.stream));
    if (gid) {
      TIT_CUDA_OK(c, c.mg.gid.ensure(c.cap_n * 8));
      TIT_CUDA_OK(c, cudaMemcpyAsync(c.mg.gid.p, gid, n_owned * 8, cudaMemcpyHostToDevice, c.stream));
    }
  }
  if (c.vt->mg_replace_owned(c, n_owned, c.mg.stageA.as<double>(), c.mg.stageB.as<double>())) return 1;
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  TITGPU_LEAVE()
}
int titgpu_mg_stats(titgpu_ctx* h, unsigned long long* exchanges, unsigned long long* migrated) {
  if (!h) return 1;
  if (exchanges) *exchanges = h->c.mg.exchanges;
  if (migrated) *migrated = h->c.mg.migrated;
  return 0;
}

int titgpu_neighbors(titgpu_ctx* h, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz) {
  TITGPU_ENTER(true)
  if (!nnz) return fail(c, "nnz must not be null");
  const int rc = c.vt->neighbors(c, off, cols, cap, nnz);
  if (rc) return rc;
  TITGPU_LEAVE()
}
int titgpu_face_neighbors(titgpu_ctx* h, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz) {
  TITGPU_ENTER(true)
  if (!nnz) return fail(c, "nnz must not be null");
  const int rc = c.vt->face_neighbors(c, off, cols, cap, nnz);
  if (rc) return rc;
  TITGPU_LEAVE()
}
int titgpu_synchronize(titgpu_ctx* h) {
  if (!h) return 1;
  Ctx& c = h->c;
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  return 0;
}
int titgpu_profile_enable(titgpu_ctx* h, int on) {
  if (!h) return 1;
  Ctx& c = h->c;
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  c.prof_fold();
  c.prof_on = on != 0;
  return 0;
}
int titgpu_profile_reset(titgpu_ctx* h) {
  if (!h) return 1;
  Ctx& c = h->c;
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  TIT_CUDA_OK(c, cudaStreamSynchronize(c.stream));
  c.prof_fold();
  c.prof_totals.clear();
  return 0;
}
int titgpu_profile_count(titgpu_ctx* h) { return h ? int(h->c.prof_totals.size()) : 0; }
int titgpu_profile_get(titgpu_ctx* h, int i, const char** name, unsigned long long* launches, double* total_ms) {
  if (!h) return 1;
  Ctx& c = h->c;
  if (i < 0 || size_t(i) >= c.prof_totals.size()) return fail(c, "profile index out of range");
  if (name) *name = c.prof_totals[size_t(i)].name.c_str();
  if (launches) *launches = c.prof_totals[size_t(i)].count;
  if (total_ms) *total_ms = c.prof_totals[size_t(i)].ms;
  return 0;
}

// 8 independent FMA chains per thread: enough ILP to saturate the FP64 pipe.
static __global__ void k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int titgpu_measure_fp64_peak(titgpu_ctx* h, double* tflops) {
  if (!h || !tflops) return 1;
  Ctx& c = h->c;
  TIT_CUDA_OK(c, cudaSetDevice(c.device));
  int sms = 0;
  TIT_CUDA_OK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  double* out = nullptr;
  TIT_CUDA_OK(c, cudaMalloc(&out, size_t(blocks) * threads * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c.stream);
    k_fp64_peak<<<blocks, threads, 0, c.stream>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c.stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    c.launches++;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  TIT_CUDA_OK(c, cudaGetLastError());
  *tflops = 2.0 * 8.0 * double(iters) * double(blocks) * threads / (double(best) * 1e-3) / 1e12;
  return 0;
}

unsigned long long titgpu_launch_count(const titgpu_ctx* h) { return h ? h->c.launches : 0; }
void* titgpu_stream(titgpu_ctx* h) { return h ? (void*)h->c.stream : nullptr; }

}  // extern "C"
