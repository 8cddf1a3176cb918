// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// C ABI (ctypes) over the CPU restatement in oracle_sim.h / oracle_kernel.h /
// oracle_math.h. Built twice by oracle/Makefile:
//   liboracle.so       -O2 -ffp-contract=off           parity oracle
//   liboracle_fast.so  -O3 -march=x86-64-v3 -fopenmp  timed CPU baseline ("port")
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load these.
#include <cstdint>
#include <cstring>
#include <memory>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_sim.h"

using namespace orc;
namespace og = oracle_gen;

namespace {

template<int D>
SimBase* make_sim(int kid) {
  switch (kid) {
    case 0: return new Sim<D, og::K0>();
    case 1: return new Sim<D, og::K1>();
    case 2: return new Sim<D, og::K2>();
    case 3: return new Sim<D, og::K3>();
    case 4: return new Sim<D, og::K4>();
    case 5: return new Sim<D, og::K5>();
    default: return nullptr;
  }
}

template<class F>
auto with_kernel(int kid, F&& f) {
  switch (kid) {
    case 0: return f(Kernel<og::K0>{});
    case 1: return f(Kernel<og::K1>{});
    case 2: return f(Kernel<og::K2>{});
    case 3: return f(Kernel<og::K3>{});
    case 4: return f(Kernel<og::K4>{});
    default: return f(Kernel<og::K5>{});
  }
}

template<int D> Vec<D> ld(const double* p) { Vec<D> v; for (int i = 0; i < D; ++i) v[i] = p[i]; return v; }
template<int D> void st(double* p, const Vec<D>& v) { for (int i = 0; i < D; ++i) p[i] = v[i]; }

}  // namespace

extern "C" {

void* orc_create(int dim, int kernel_id, int eos_id, int integrator_id) {
  SimBase* s = dim == 2 ? make_sim<2>(kernel_id) : dim == 3 ? make_sim<3>(kernel_id) : nullptr;
  if (!s) return nullptr;
  s->prm.eos = eos_id;
  s->prm.integrator = integrator_id;
  return s;
}
void orc_destroy(void* h) { delete static_cast<SimBase*>(h); }

int orc_set_params(void* h, double g, double mu, double cs0, double rho0, double xi, double hh, double search_hint, double face_hint) {
  auto* s = static_cast<SimBase*>(h);
  s->prm.g = g; s->prm.mu = mu; s->prm.cs0 = cs0; s->prm.rho0 = rho0; s->prm.xi = xi; s->prm.h = hh;
  s->prm.search_hint = search_hint; s->prm.face_hint = face_hint;
  return 0;
}
// 1: the pair sums run over unordered pairs and update both particles (the reference's own loop
// structure, block-coloured: SimBase Sim::for_each_pair); 0 (default): gather form with a fixed order.
int orc_set_symmetric(void* h, int on) { static_cast<SimBase*>(h)->prm.symmetric = on != 0; return 0; }
int orc_set_surface(void* h, const double* v, size_t nv, const uint64_t* f, size_t nf, const double* cv, size_t ncv, const uint64_t* cf, size_t ncf) {
  return static_cast<SimBase*>(h)->set_surface(v, nv, f, nf, cv, ncv, cf, ncf);
}
int orc_resize(void* h, size_t n_fluid, size_t n_fixed) { static_cast<SimBase*>(h)->resize(n_fluid, n_fixed); return 0; }
int orc_upload(void* h, const char* field, const double* data) {
  auto* s = static_cast<SimBase*>(h);
  double* p; int w;
  if (s->field(field, &p, &w)) return 1;
  if (s->n()) std::memcpy(p, data, s->n() * w * sizeof(double));
  return 0;
}
int orc_download(void* h, const char* field, double* data) {
  auto* s = static_cast<SimBase*>(h);
  double* p; int w;
  if (s->field(field, &p, &w)) return 1;
  if (s->n()) std::memcpy(data, p, s->n() * w * sizeof(double));
  return 0;
}
int orc_initialize(void* h) { static_cast<SimBase*>(h)->initialize(); return 0; }
int orc_prepare(void* h) { static_cast<SimBase*>(h)->prepare(); return 0; }
int orc_rhs_only(void* h) { static_cast<SimBase*>(h)->rhs_only(); return 0; }
// FluidEquations::post_integrate alone (fluid_equations.hpp:315-321) on the current state.
int orc_post_only(void* h) { static_cast<SimBase*>(h)->post_only(); return 0; }
// Seconds per phase since the last call with reset != 0 (see SimBase::phase_s).
int orc_phase_times(void* h, double* out8, int reset) {
  auto* s = static_cast<SimBase*>(h);
  for (int i = 0; i < 8; ++i) { out8[i] = s->phase_s[i]; if (reset) s->phase_s[i] = 0.0; }
  return 0;
}
// Branch counters of the last post_integrate (see SimBase::stats).
int orc_stats(void* h, long long* out8) { for (int i = 0; i < 8; ++i) out8[i] = static_cast<SimBase*>(h)->stats[i]; return 0; }
int orc_step(void* h, int nsteps, double* dt_last) {
  auto* s = static_cast<SimBase*>(h);
  double dt = 0;
  for (int i = 0; i < nsteps; ++i) dt = s->step();
  if (dt_last) *dt_last = dt;
  return 0;
}
// CSR of sorted neighbour rows; call with cols == NULL to get nnz only.
int orc_neighbors(void* h, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz) {
  auto* s = static_cast<SimBase*>(h);
  std::vector<uint64_t> o, c;
  s->neighbors(o, c);
  *nnz = c.size();
  if (!cols) return 0;
  if (cap < c.size()) return 2;
  std::memcpy(off, o.data(), o.size() * 8);
  std::memcpy(cols, c.data(), c.size() * 8);
  return 0;
}
int orc_face_neighbors(void* h, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz) {
  auto* s = static_cast<SimBase*>(h);
  std::vector<uint64_t> o, c;
  s->face_neighbors(o, c);
  *nnz = c.size();
  if (!cols) return 0;
  if (cap < c.size()) return 2;
  std::memcpy(off, o.data(), o.size() * 8);
  std::memcpy(cols, c.data(), c.size() * 8);
  return 0;
}
int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// ---- kernel layer, for the reference's known-answer tests -----------------
double orc_tiny() { return tiny; }
double orc_kernel_radius(int kid, double h) { return with_kernel(kid, [&](auto k) { return decltype(k)::radius(h); }); }
double orc_kernel_weight(int kid, int dim) {
  return with_kernel(kid, [&](auto k) { using KK = decltype(k); return dim == 1 ? KK::template weight<1>() : dim == 2 ? KK::template weight<2>() : KK::template weight<3>(); });
}
double orc_kernel_unit_value(int kid, double q) { return with_kernel(kid, [&](auto k) { (void)k; switch (kid) { case 0: return og::K0::unit_value(q); case 1: return og::K1::unit_value(q); case 2: return og::K2::unit_value(q); case 3: return og::K3::unit_value(q); case 4: return og::K4::unit_value(q); default: return og::K5::unit_value(q); } }); }
double orc_kernel_unit_deriv(int kid, double q) { switch (kid) { case 0: return og::K0::unit_deriv(q); case 1: return og::K1::unit_deriv(q); case 2: return og::K2::unit_deriv(q); case 3: return og::K3::unit_deriv(q); case 4: return og::K4::unit_deriv(q); default: return og::K5::unit_deriv(q); } }
double orc_kernel_value(int kid, int dim, const double* x, double h) {
  return with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 1) return KK::template value<1>(ld<1>(x), h);
    if (dim == 2) return KK::template value<2>(ld<2>(x), h);
    return KK::template value<3>(ld<3>(x), h);
  });
}
void orc_kernel_grad(int kid, int dim, const double* x, double h, double* out) {
  with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 1) st<1>(out, KK::template grad<1>(ld<1>(x), h));
    else if (dim == 2) st<2>(out, KK::template grad<2>(ld<2>(x), h));
    else st<3>(out, KK::template grad<3>(ld<3>(x), h));
    return 0;
  });
}
double orc_kernel_width_deriv(int kid, int dim, const double* x, double h) {
  return with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 1) return KK::template width_deriv<1>(ld<1>(x), h);
    if (dim == 2) return KK::template width_deriv<2>(ld<2>(x), h);
    return KK::template width_deriv<3>(ld<3>(x), h);
  });
}
void orc_kernel_antigrad(int kid, int dim, const double* x, double h, double* out) {
  with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 1) st<1>(out, KK::template antigrad<1>(ld<1>(x), h));
    else if (dim == 2) st<2>(out, KK::template antigrad<2>(ld<2>(x), h));
    else st<3>(out, KK::template antigrad<3>(ld<3>(x), h));
    return 0;
  });
}
// face: dim vertices of dim doubles each (segment in 2-D, triangle in 3-D).
void orc_kernel_flux(int kid, int dim, const double* face, const double* x, double h, double* out) {
  with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 2) st<2>(out, KK::flux(Segment{ld<2>(face), ld<2>(face + 2)}, ld<2>(x), h));
    else st<3>(out, KK::flux(Triangle{ld<3>(face), ld<3>(face + 3), ld<3>(face + 6)}, ld<3>(x), h));
    return 0;
  });
}
double orc_kernel_antigrad_flux(int kid, int dim, const double* face, const double* x, double h) {
  return with_kernel(kid, [&](auto k) {
    using KK = decltype(k);
    if (dim == 2) return KK::antigrad_flux(Segment{ld<2>(face), ld<2>(face + 2)}, ld<2>(x), h);
    return KK::antigrad_flux(Triangle{ld<3>(face), ld<3>(face + 3), ld<3>(face + 6)}, ld<3>(x), h);
  });
}

// Batched variants (n points, packed) so that the quadrature-based known-answer
// tests of sph/kernel.test.cpp can be restated without per-point call overhead.
void orc_kernel_value_n(int kid, int dim, const double* x, size_t n, double h, double* out) {
  for (size_t i = 0; i < n; ++i) out[i] = orc_kernel_value(kid, dim, x + i * dim, h);
}
void orc_kernel_antigrad_n(int kid, int dim, const double* x, size_t n, double h, double* out) {
  for (size_t i = 0; i < n; ++i) orc_kernel_antigrad(kid, dim, x + i * dim, h, out + i * dim);
}

// ---- geometry layer ----------------------------------------------------------
void orc_segment_clamp(const double* seg, const double* p, double* out) { st<2>(out, Segment{ld<2>(seg), ld<2>(seg + 2)}.clamp(ld<2>(p))); }
void orc_triangle_clamp(const double* tri, const double* p, double* out) { st<3>(out, Triangle{ld<3>(tri), ld<3>(tri + 3), ld<3>(tri + 6)}.clamp(ld<3>(p))); }
int orc_face_intersects(int dim, const double* face, const double* c, double radius) {
  if (dim == 2) return Segment{ld<2>(face), ld<2>(face + 2)}.intersects(BSphere<2>{ld<2>(c), radius});
  return Triangle{ld<3>(face), ld<3>(face + 3), ld<3>(face + 6)}.intersects(BSphere<3>{ld<3>(c), radius});
}
double orc_winding(int dim, const double* verts, size_t nv, const uint64_t* faces, size_t nf, const double* p) {
  double w = 0;
  (void)nv;
  for (size_t f = 0; f < nf; ++f) {
    if (dim == 2) w += Segment{ld<2>(verts + 2 * faces[2 * f]), ld<2>(verts + 2 * faces[2 * f + 1])}.winding_number(ld<2>(p));
    else w += Triangle{ld<3>(verts + 3 * faces[3 * f]), ld<3>(verts + 3 * faces[3 * f + 1]), ld<3>(verts + 3 * faces[3 * f + 2])}.winding_number(ld<3>(p));
  }
  return w;
}
// Grid cell math (geom/grid.hpp): box lo/hi, hint -> num cells, extents; and the
// inclusive cell range overlapping a query box.
int orc_grid(int dim, const double* lo, const double* hi, double hint, uint64_t* num, double* ext) {
  auto run = [&](auto g) {
    constexpr int D = decltype(g)::value;
    Grid<D> grid;
    grid.box.lo = ld<D>(lo); grid.box.hi = ld<D>(hi);
    grid.set_cell_extents(hint);
    for (int i = 0; i < D; ++i) { num[i] = grid.num[i]; ext[i] = grid.ext[i]; }
    return 0;
  };
  return dim == 2 ? run(std::integral_constant<int, 2>{}) : run(std::integral_constant<int, 3>{});
}
int orc_grid_cells_intersecting(int dim, const double* lo, const double* hi, double hint, const double* qlo, const double* qhi, uint64_t* clo, uint64_t* chi) {
  auto run = [&](auto g) {
    constexpr int D = decltype(g)::value;
    Grid<D> grid;
    grid.box.lo = ld<D>(lo); grid.box.hi = ld<D>(hi);
    grid.set_cell_extents(hint);
    BBox<D> q; q.lo = ld<D>(qlo); q.hi = ld<D>(qhi);
    std::array<size_t, D> a, b;
    if (!grid.cells_intersecting(q, a, b)) return 1;
    for (int i = 0; i < D; ++i) { clo[i] = a[i]; chi[i] = b[i]; }
    return 0;
  };
  return dim == 2 ? run(std::integral_constant<int, 2>{}) : run(std::integral_constant<int, 3>{});
}
// Neighbour rows through the K-d tree index (kd_tree_search.hpp) as ParticleMesh::search_
// builds them (particle_mesh.hpp:137-147: every point's own sphere, rows sorted).
int orc_kdtree_neighbors(int dim, const double* pts, size_t n, double radius, uint64_t* off, uint64_t* cols, size_t cap, size_t* nnz) {
  auto run = [&](auto g) {
    constexpr int D = decltype(g)::value;
    std::vector<Vec<D>> p(n);
    for (size_t i = 0; i < n; ++i) p[i] = ld<D>(pts + i * D);
    const KDTree<D> tree(p);
    std::vector<uint64_t> row;
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
      row.clear();
      BSphere<D> s; s.c = p[i]; s.r = radius;
      tree.search(s, [&](size_t j) { row.push_back(j); });
      std::sort(row.begin(), row.end());
      if (cols && total + row.size() <= cap) std::memcpy(cols + total, row.data(), row.size() * 8);
      if (off) off[i] = total;
      total += row.size();
    }
    if (off) off[n] = total;
    *nnz = total;
    return cols && total > cap ? 2 : 0;
  };
  return dim == 2 ? run(std::integral_constant<int, 2>{}) : run(std::integral_constant<int, 3>{});
}
int orc_lu_inverse(int dim, const double* A, double* inv) {
  if (dim == 2) { Mat<2> a, r; std::memcpy(&a, A, sizeof a); if (!lu_inverse(a, r)) return 1; std::memcpy(inv, &r, sizeof r); return 0; }
  Mat<3> a, r; std::memcpy(&a, A, sizeof a); if (!lu_inverse(a, r)) return 1; std::memcpy(inv, &r, sizeof r); return 0;
}

}  // extern "C"
