// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// CPU restatement of the reference WCSPH particle step in gather form
// (SURVEY.md Appendix A). Follows, under /root/reference/source/tit/:
//   sph/fluid_equations.hpp:79-524   initialize / prepare / compute_gamma /
//                                    setup_boundary / compute_time_step /
//                                    compute_continuity / compute_momentum /
//                                    apply_shifts / apply_free_surface_correction
//   sph/time_integrator.hpp:32-228   Euler, Verlet, SSPRK2/3
//   sph/equation_of_state.hpp:19-122 Tait / linear Tait
//   sph/particle_mesh.hpp:124-162    adjacency: self included, sorted, `<=`
//   geom/search/grid_search.hpp:45-102, geom/grid.hpp:90-168 cell list
//   geom/face_search/grid_face_search.hpp:43-112 face cell list
//
// Parity status: the kernel / geometry layers are pinned by the reference's
// known-answer tests (tests/test_oracle_kernels.py, test_oracle_geometry.py).
// FluidEquations, ParticleMesh and the integrators have NO reference unit test
// and the reference cannot be built here (C++26, GCC 16): they are pinned by an
// independent numpy transcription of the reference's lambdas in their original
// symmetric scatter form plus known physical answers (tests/test_oracle_physics.py;
// DESIGN.md section 5). The 3-D dam break has no counterpart in the reference.
//
// The pair loops of the reference are symmetric (update a and b per unordered
// pair); the gather form used here is algebraically and bitwise term-identical
// (Psi_ab = Psi_ba, Pi_ab = Pi_ba, P_ab = P_ba, grad W_ab = -grad W_ba) and
// sums each particle's terms in ascending neighbour index. prm.symmetric = 1
// switches the pair sums to the reference's own structure (for_each_pair): the
// CPU baseline of bench.py uses that, the parity tests never do.
#pragma once

#include <chrono>
#include <cstdio>
#include <string>

#include "oracle_kernel.h"

namespace orc {

struct Params {
  double g = 9.81, mu = 1e-3, cs0 = 1.0, rho0 = 1000.0, xi = 7.0, h = 1.0;
  double search_hint = 0.0, face_hint = 0.0;
  int eos = 0;         // 0 Tait, 1 linear Tait
  int integrator = 3;  // 0 symplectic Euler, 1 velocity Verlet, 2 SSPRK2, 3 SSPRK3
  int symmetric = 0;   // 1: pair sums over unordered pairs, both particles updated per pair (see for_each_pair)
};

struct SimBase {
  virtual ~SimBase() = default;
  Params prm;
  std::size_t nf = 0, nx = 0;  // fluid, fixed
  std::string err;
  std::size_t n() const { return nf + nx; }
  virtual int dim() const = 0;
  virtual int set_surface(const double* v, std::size_t nv, const std::uint64_t* f, std::size_t nfaces, const double* cv, std::size_t ncv, const std::uint64_t* cf, std::size_t ncf) = 0;
  virtual void resize(std::size_t n_fluid, std::size_t n_fixed) = 0;
  virtual int field(const char* name, double** ptr, int* width) = 0;
  virtual void initialize() = 0;
  virtual void prepare() = 0;
  virtual void rhs_only() = 0;
  virtual void post_only() = 0;
  virtual double step() = 0;
  // Branch counters of the last post_integrate (tests assert that the interesting branches
  // of fluid_equations.hpp:366-511 were exercised): [0] LU of L^T failed -> identity,
  // [1] particles on the free surface after the visibility test + splash rule, [2] of those
  // by the splash rule alone, [3] near-surface particles (phi scaled), [4] shifted particles,
  // [5] shifted with the velocity correction skipped (|gamma - 1| > tiny), [6] free-surface
  // corrections applied (ratio <= 0.99), [7] candidates of the correction (phi != 1).
  long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // Wall-clock seconds per phase, accumulated since the last reset (CPU-baseline breakdown):
  // [0] search (particle + face adjacency), [1] compute_gamma, [2] setup_boundary,
  // [3] continuity + momentum (face terms and pair sums), [4] integrator update / lincomb / dt,
  // [5] apply_shifts, [6] free-surface correction.
  double phase_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  struct PhaseTimer {
    double& acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit PhaseTimer(double& a) : acc(a) {}
    ~PhaseTimer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  };
  virtual void neighbors(std::vector<std::uint64_t>& off, std::vector<std::uint64_t>& cols) = 0;
  virtual void face_neighbors(std::vector<std::uint64_t>& off, std::vector<std::uint64_t>& cols) = 0;
};

template<int D, class KG>
struct Sim final : SimBase {
  using K = Kernel<KG>;
  using V = Vec<D>;
  using M = Mat<D>;
  using Face = typename FaceOf<D>::type;

  // Varying fields (fluid_equations.hpp:41-48), packed.
  std::vector<double> m, gamma, rho, drho_dt, p, cs, phi, rho_raw;
  std::vector<V> grad_gamma, grad_rho, v, dv_dt, r, dr, N;
  std::vector<M> grad_v, L;

  Surface<D> domain, containment;

  // Adjacency (CSR, sorted ascending, self included).
  std::vector<std::uint32_t> nb_off, nb;
  std::vector<std::uint32_t> fc_off, fc;

  // Static face index (the surface never moves; the reference rebuilds it per
  // prepare with identical result).
  Grid<D> fgrid;
  std::vector<std::uint32_t> fcell_off, fcell_faces;
  bool fgrid_ok = false;

  static constexpr double CFL = 0.4, C_force = 0.25, C_visc = 0.125, C_shift = 0.2;
  static constexpr double phi_max = 1.0;
  static constexpr double phi_min = std::numeric_limits<double>::min();

  int dim() const override { return D; }

  // ---- EOS (equation_of_state.hpp) ----
  double eos_p(double rho_) const {
    if (prm.eos == 1) return pow2(prm.cs0) * (rho_ - prm.rho0);
    const double B = prm.rho0 * pow2(prm.cs0) / prm.xi;
    return B * (std::pow(rho_ / prm.rho0, prm.xi) - 1.0);
  }
  double eos_cs(double rho_) const {
    if (prm.eos == 1) return prm.cs0;
    return prm.cs0 * std::pow(rho_ / prm.rho0, (prm.xi - 1.0) / 2.0);
  }
  double eos_H(double rho_) const {
    if (prm.eos == 1) return pow2(prm.cs0) * std::log(rho_ / prm.rho0);
    const double x1 = prm.xi - 1.0;
    return pow2(prm.cs0) * (std::pow(rho_ / prm.rho0, x1) - 1.0) / x1;
  }
  double eos_rho_from_H(double H) const {
    if (prm.eos == 1) return prm.rho0 * std::exp(H / pow2(prm.cs0));
    const double x1 = prm.xi - 1.0;
    return prm.rho0 * std::pow(1.0 + x1 * H / pow2(prm.cs0), 1.0 / x1);
  }

  // ---- setup ----
  int set_surface(const double* vv, std::size_t nv, const std::uint64_t* f, std::size_t nfaces, const double* cv, std::size_t ncv, const std::uint64_t* cf, std::size_t ncf) override {
    auto fill = [](Surface<D>& s, const double* vp, std::size_t nvp, const std::uint64_t* fp, std::size_t nfp) {
      s.verts.resize(nvp);
      for (std::size_t i = 0; i < nvp; ++i)
        for (int d = 0; d < D; ++d) s.verts[i][d] = vp[i * D + d];
      s.faces.resize(nfp);
      for (std::size_t i = 0; i < nfp; ++i)
        for (int d = 0; d < D; ++d) s.faces[i][d] = fp[i * D + d];
    };
    fill(domain, vv, nv, f, nfaces);
    fill(containment, cv, ncv, cf, ncf);
    fgrid_ok = false;
    return 0;
  }

  void resize(std::size_t n_fluid, std::size_t n_fixed) override {
    nf = n_fluid;
    nx = n_fixed;
    const std::size_t nn = n();
    for (auto* s : {&m, &gamma, &rho, &drho_dt, &p, &cs, &phi, &rho_raw}) s->assign(nn, 0.0);
    for (auto* s : {&grad_gamma, &grad_rho, &v, &dv_dt, &r, &dr, &N}) s->assign(nn, V{});
    for (auto* s : {&grad_v, &L}) s->assign(nn, M{});
  }

  int field(const char* name, double** ptr, int* width) override {
    const std::string s(name);
#define ORC_S(f) if (s == #f) { *ptr = f.data(); *width = 1; return 0; }
#define ORC_V(f) if (s == #f) { *ptr = f.empty() ? nullptr : f[0].data(); *width = D; return 0; }
#define ORC_M(f) if (s == #f) { *ptr = f.empty() ? nullptr : f[0].data(); *width = D * D; return 0; }
    ORC_S(m) ORC_S(gamma) ORC_S(rho) ORC_S(drho_dt) ORC_S(p) ORC_S(cs) ORC_S(phi) ORC_S(rho_raw)
    ORC_V(grad_gamma) ORC_V(grad_rho) ORC_V(v) ORC_V(dv_dt) ORC_V(r) ORC_V(dr) ORC_V(N)
    ORC_M(grad_v) ORC_M(L)
#undef ORC_S
#undef ORC_V
#undef ORC_M
    return 1;
  }

  bool is_fluid(std::size_t a) const { return a < nf; }
  double radius() const { return K::radius(prm.h); }

  // ---- neighbour search (particle_mesh.hpp:124-162) ----
  void build_face_grid() {
    // grid_face_search.hpp:43-88.
    fgrid_ok = true;
    fcell_off.clear();
    fcell_faces.clear();
    if (domain.faces.empty()) return;
    const double hint = prm.face_hint > 0 ? prm.face_hint : prm.h;
    BBox<D> box(domain.verts[0]);
    for (const auto& q : domain.verts) box.expand(q);
    box.grow(hint / 2);
    fgrid.box = box;
    fgrid.set_cell_extents(hint);
    const std::size_t nc = fgrid.flat_num();
    std::vector<std::uint32_t> cnt(nc + 1, 0);
    auto for_cells = [&](std::size_t f, auto&& fn) {
      std::array<std::size_t, D> lo, hi;
      if (!fgrid.cells_intersecting(domain.face(f).box(), lo, hi)) return;
      std::array<std::size_t, D> c = lo;
      for (;;) {
        fn(fgrid.flatten(c));
        int d = D - 1;
        while (d >= 0 && ++c[d] > hi[d]) { c[d] = lo[d]; --d; }
        if (d < 0) break;
      }
    };
    for (std::size_t f = 0; f < domain.faces.size(); ++f) for_cells(f, [&](std::size_t c) { cnt[c + 1]++; });
    for (std::size_t c = 0; c < nc; ++c) cnt[c + 1] += cnt[c];
    fcell_off = cnt;
    fcell_faces.resize(cnt[nc]);
    std::vector<std::uint32_t> pos(cnt.begin(), cnt.end() - 1);
    for (std::size_t f = 0; f < domain.faces.size(); ++f) for_cells(f, [&](std::size_t c) { fcell_faces[pos[c]++] = std::uint32_t(f); });
  }

  void search() {
    const std::size_t nn = n();
    const double rad = radius();
    const double hint = prm.search_hint > 0 ? prm.search_hint : prm.h;
    // GridIndex (grid_search.hpp:45-85): bbox grown by hint/2, stretched cells.
    Grid<D> grid;
    std::vector<std::uint32_t> cell_off, cell_pts;
    if (nn > 0) {
      BBox<D> box(r[0]);
      for (std::size_t a = 0; a < nn; ++a) box.expand(r[a]);
      box.grow(hint / 2);
      grid.box = box;
      grid.set_cell_extents(hint);
      const std::size_t nc = grid.flat_num();
      cell_off.assign(nc + 1, 0);
      std::vector<std::uint32_t> pc(nn);
      for (std::size_t a = 0; a < nn; ++a) { pc[a] = std::uint32_t(grid.flatten(grid.cell_index(r[a]))); cell_off[pc[a] + 1]++; }
      for (std::size_t c = 0; c < nc; ++c) cell_off[c + 1] += cell_off[c];
      cell_pts.resize(nn);
      std::vector<std::uint32_t> pos(cell_off.begin(), cell_off.end() - 1);
      for (std::size_t a = 0; a < nn; ++a) cell_pts[pos[pc[a]]++] = std::uint32_t(a);  // ascending within a cell
    }
    // Per-particle sphere query, two passes (count, fill); rows sorted ascending.
    auto query = [&](std::size_t a, auto&& emit) {
      const BSphere<D> sph{r[a], rad};
      std::array<std::size_t, D> lo, hi;
      if (!grid.cells_intersecting(sph.box(), lo, hi)) return;
      std::array<std::size_t, D> c = lo;
      for (;;) {
        const std::size_t fc_ = grid.flatten(c);
        for (std::uint32_t i = cell_off[fc_]; i < cell_off[fc_ + 1]; ++i)
          if (sph.contains(r[cell_pts[i]])) emit(cell_pts[i]);
        int d = D - 1;
        while (d >= 0 && ++c[d] > hi[d]) { c[d] = lo[d]; --d; }
        if (d < 0) break;
      }
    };
    nb_off.assign(nn + 1, 0);
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) { std::uint32_t k = 0; query(a, [&](std::uint32_t) { ++k; }); nb_off[a + 1] = k; }
    for (std::size_t a = 0; a < nn; ++a) nb_off[a + 1] += nb_off[a];
    nb.resize(nb_off[nn]);
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) {
      std::uint32_t k = nb_off[a];
      query(a, [&](std::uint32_t b) { nb[k++] = b; });
      std::sort(nb.begin() + nb_off[a], nb.begin() + nb_off[a + 1]);
    }

    // Faces (grid_face_search.hpp:91-112): dedupe + exact intersects().
    if (!fgrid_ok) build_face_grid();
    fc_off.assign(nn + 1, 0);
    std::vector<std::vector<std::uint32_t>> rows(nn);
    if (!domain.faces.empty()) {
#pragma omp parallel for schedule(dynamic, 256)
      for (std::size_t a = 0; a < nn; ++a) {
        const BSphere<D> sph{r[a], rad};
        std::array<std::size_t, D> lo, hi;
        if (!fgrid.cells_intersecting(sph.box(), lo, hi)) continue;
        auto& row = rows[a];
        std::array<std::size_t, D> c = lo;
        for (;;) {
          const std::size_t fcl = fgrid.flatten(c);
          for (std::uint32_t i = fcell_off[fcl]; i < fcell_off[fcl + 1]; ++i) {
            row.push_back(fcell_faces[i]);
          }
          int d = D - 1;
          while (d >= 0 && ++c[d] > hi[d]) { c[d] = lo[d]; --d; }
          if (d < 0) break;
        }
        // A face is listed in every cell it overlaps: dedupe (the reference uses
        // per-thread visited marks), then apply the exact test.
        std::sort(row.begin(), row.end());
        row.erase(std::unique(row.begin(), row.end()), row.end());
        row.erase(std::remove_if(row.begin(), row.end(), [&](std::uint32_t f) { return !domain.face(f).intersects(sph); }), row.end());
      }
    }
    for (std::size_t a = 0; a < nn; ++a) fc_off[a + 1] = fc_off[a] + std::uint32_t(rows[a].size());
    fc.resize(fc_off[nn]);
    for (std::size_t a = 0; a < nn; ++a) std::copy(rows[a].begin(), rows[a].end(), fc.begin() + fc_off[a]);
  }

  // Face-vertex averages (field.hpp:59-65; face vertex k <-> fixed particle k).
  template<class T> T favg(const std::vector<T>& f, std::uint32_t face) const {
    const auto& fv = domain.faces[face];
    if constexpr (std::is_same_v<T, double>) {
      double s = f[nf + fv[0]];
      for (int k = 1; k < D; ++k) s += f[nf + fv[k]];
      return s / double(D);
    } else {
      T s = f[nf + fv[0]];
      for (int k = 1; k < D; ++k) s += f[nf + fv[k]];
      return s / double(D);
    }
  }

  // ---- fluid_equations.hpp:171-193 ----
  void compute_gamma() {
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nn; ++a) {
      V gg{};
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) gg += K::flux(domain.face(fc[i]), r[a], h);
      grad_gamma[a] = gg;
      double ga = containment.contains(r[a]) ? 1.0 : 0.0;
      const double ng = norm(gg);
      if (!is_tiny(ng)) {
        const V n_a = gg / ng;
        const V r_a = r[a] + ((2.0 * ga - 1.0) * n_a) * pow2(h);
        for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) ga -= K::antigrad_flux(domain.face(fc[i]), r_a, h);
      }
      gamma[a] = ga;
    }
  }

  // ---- fluid_equations.hpp:122-164 ----
  void setup_boundary() {
    const double h = prm.h;
#pragma omp parallel for schedule(dynamic, 64)
    for (std::size_t e = nf; e < n(); ++e) {
      v[e] = V{};
      double S_e = 0.0, H_e = 0.0;
      const V n_e = normalize(grad_gamma[e]);
      for (std::uint32_t i = nb_off[e]; i < nb_off[e + 1]; ++i) {
        const std::size_t b = nb[i];
        if (!is_fluid(b)) continue;
        const double V_b = m[b] / rho[b];
        const V r_be = r[b] - r[e];
        const double W_be = K::template value<D>(r_be, h);
        const double H_b = eos_H(rho[b]);
        S_e += V_b * W_be;
        H_e += V_b * (H_b + prm.g * dot(r_be, n_e) * n_e[1]) * W_be;
      }
      rho[e] = eos_rho_from_H(is_tiny(S_e) ? 0.0 : H_e / S_e);
    }
  }

  void prepare() override {
    { PhaseTimer t(phase_s[0]); search(); }
    { PhaseTimer t(phase_s[1]); compute_gamma(); }
    { PhaseTimer t(phase_s[2]); setup_boundary(); }
  }

  void initialize() override {
    search();
    compute_gamma();
    for (std::size_t a = nf; a < n(); ++a) m[a] *= gamma[a];
  }

  // ---- fluid_equations.hpp:199-222 ----
  double compute_time_step() const {
    double dt = std::numeric_limits<double>::max();
    const double h = prm.h;
    for (std::size_t a = 0; a < nf; ++a) {
      const double dt_ac = CFL * h / (eos_cs(rho[a]) + norm(v[a]));
      const double dt_visc = C_visc * pow2(h) * rho[a] / prm.mu;
      const double dt_force = C_force * std::sqrt(h / std::max(norm(dv_dt[a]), prm.g));
      dt = std::min({dt, dt_ac, dt_visc, dt_force});
    }
    return dt;
  }

  // ---- symmetric pair loops ----
  // The reference evaluates every unordered pair ONCE and updates both particles
  // (fluid_equations.hpp:249-259, 293-304, 351-365), made race-free by a two-level
  // partition of the particles into blocks (particle_mesh.hpp:165-241): pairs inside a
  // block run in parallel over the blocks, the pairs between blocks in a later pass. The
  // same scheme with slabs along x at least one support radius thick: a pair joins
  // particles of one slab or of two adjacent slabs, so three passes cover every pair
  // exactly once without two threads touching one particle - slab-internal pairs (all
  // slabs in parallel), pairs (i, i + 1) for even i, then for odd i. Used by the CPU
  // baseline of bench.py (prm.symmetric = 1); the parity oracle keeps the gather form,
  // whose sums have a fixed order. tests/test_oracle_physics.py checks that the two agree.
  std::vector<std::uint32_t> blk_of, blk_off, blk_items;
  void build_blocks() {
    const std::size_t nn = n();
    double lo = 1e300, hi = -1e300;
    for (std::size_t a = 0; a < nn; ++a) { lo = std::min(lo, r[a][0]); hi = std::max(hi, r[a][0]); }
    const double width = radius() * 1.0001;
    const std::size_t nb_ = std::size_t(std::max(1.0, std::floor((hi - lo) / width))) + 1;
    blk_of.assign(nn, 0);
    blk_off.assign(nb_ + 1, 0);
    for (std::size_t a = 0; a < nn; ++a) {
      blk_of[a] = std::uint32_t(std::min<double>(double(nb_ - 1), std::floor((r[a][0] - lo) / width)));
      blk_off[blk_of[a] + 1]++;
    }
    for (std::size_t i = 0; i < nb_; ++i) blk_off[i + 1] += blk_off[i];
    blk_items.resize(nn);
    std::vector<std::uint32_t> pos(blk_off.begin(), blk_off.end() - 1);
    for (std::size_t a = 0; a < nn; ++a) blk_items[pos[blk_of[a]]++] = std::uint32_t(a);
  }
  // fn(a, b) once per unordered pair of distinct neighbours.
  template<class F> void for_each_pair(F&& fn) {
    build_blocks();
    const long nblk = long(blk_off.size()) - 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < nblk; ++i)
      for (std::uint32_t k = blk_off[i]; k < blk_off[i + 1]; ++k) {
        const std::size_t a = blk_items[k];
        for (std::uint32_t j = nb_off[a]; j < nb_off[a + 1]; ++j) {
          const std::size_t b = nb[j];
          if (b > a && blk_of[b] == std::uint32_t(i)) fn(a, b);
        }
      }
    for (long parity = 0; parity < 2; ++parity) {
#pragma omp parallel for schedule(dynamic, 1)
      for (long i = parity; i < nblk - 1; i += 2)
        for (std::uint32_t k = blk_off[i]; k < blk_off[i + 1]; ++k) {
          const std::size_t a = blk_items[k];
          for (std::uint32_t j = nb_off[a]; j < nb_off[a + 1]; ++j) {
            const std::size_t b = nb[j];
            if (blk_of[b] == std::uint32_t(i + 1)) fn(a, b);
          }
        }
    }
  }
  void compute_continuity_symmetric() {
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) cs[a] = eos_cs(rho[a]);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      double acc = 0.0;
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        acc -= favg(rho, s) * dot(v[a] - favg(v, s), gg) / gamma[a];
      }
      drho_dt[a] = acc;
    }
    for_each_pair([&](std::size_t a, std::size_t b) {
      const V r_ab = r[a] - r[b];
      const V gW = K::template grad<D>(r_ab, h);
      const double cs_ab = std::max(cs[a], cs[b]);
      const V Psi = (cs_ab * (rho[a] - rho[b]) * r_ab) / norm(r_ab);  // Psi_ab = Psi_ba
      const V v_ab = v[a] - v[b];
      if (a < nf) drho_dt[a] += m[b] / gamma[a] * dot(v_ab + Psi / rho[b], gW);
      if (b < nf) drho_dt[b] += m[a] / gamma[b] * dot(v_ab - Psi / rho[a], gW);  // (v_ba + Psi / rho_a) . grad W_ba
    });
  }
  void compute_momentum_symmetric() {
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) p[a] = eos_p(rho[a]);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      V acc{};
      acc[1] = -prm.g;
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        const double rho_s = favg(rho, s), p_s = favg(p, s);
        const double P_as = rho_s * (p[a] / pow2(rho[a]) + p_s / pow2(rho_s));
        const V n_s = normalize(gg);
        const V v_as = v[a] - favg(v, s);
        const V t_as = normalize(v_as - dot(v_as, n_s) * n_s);
        const double dr_as = std::max(h / 2.0, dot(r[a] - favg(r, s), n_s));
        const V Pi_as = (2.0 * prm.mu / (rho[a] * dr_as) * dot(v_as, t_as)) * t_as;
        acc += (P_as * gg - Pi_as * norm(gg)) / gamma[a];
      }
      dv_dt[a] = acc;
    }
    for_each_pair([&](std::size_t a, std::size_t b) {
      const V r_ab = r[a] - r[b];
      const V gW = K::template grad<D>(r_ab, h);
      const double P_ab = p[a] / pow2(rho[a]) + p[b] / pow2(rho[b]);
      const double Pi_ab = 2.0 * prm.mu * dot(v[a] - v[b], r_ab) / (rho[a] * rho[b] * norm2(r_ab));
      if (a < nf) dv_dt[a] += (m[b] / gamma[a] * (Pi_ab - P_ab)) * gW;
      if (b < nf) dv_dt[b] -= (m[a] / gamma[b] * (Pi_ab - P_ab)) * gW;
    });
  }
  // The sums of apply_shifts over unordered pairs: raw Na -> dr, La -> L, gv -> grad_v, gr -> grad_rho.
  void shift_sums_symmetric() {
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nn; ++a) {
      V Na{}, gr{};
      M La{}, gv{};
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        Na -= gg / gamma[a];
        La -= outer(favg(r, s) - r[a], gg) / gamma[a];
        gv -= outer(favg(v, s) - v[a], gg) / gamma[a];
        gr -= ((favg(rho, s) - rho[a]) * gg) / gamma[a];
      }
      dr[a] = Na; L[a] = La; grad_v[a] = gv; grad_rho[a] = gr;
    }
    for_each_pair([&](std::size_t a, std::size_t b) {
      const V gW = K::template grad<D>(r[a] - r[b], h);
      const M o_r = outer(r[b] - r[a], gW), o_v = outer(v[b] - v[a], gW);
      const V g_rho = (rho[b] - rho[a]) * gW;
      const double ca = m[b] / rho[b] / gamma[a], cb = m[a] / rho[a] / gamma[b];
      dr[a] += ca * gW; L[a] += ca * o_r; grad_v[a] += ca * o_v; grad_rho[a] += ca * g_rho;
      // b's terms: grad W_ba = -grad W_ab, r_a - r_b = -(r_b - r_a), ...: the outer products and g_rho keep their sign
      dr[b] -= cb * gW; L[b] += cb * o_r; grad_v[b] += cb * o_v; grad_rho[b] += cb * g_rho;
    });
  }

  // ---- fluid_equations.hpp:232-260 ----
  void compute_continuity() {
    if (prm.symmetric) return compute_continuity_symmetric();
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) cs[a] = eos_cs(rho[a]);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      double acc = 0.0;
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        acc -= favg(rho, s) * dot(v[a] - favg(v, s), gg) / gamma[a];
      }
      for (std::uint32_t i = nb_off[a]; i < nb_off[a + 1]; ++i) {
        const std::size_t b = nb[i];
        if (b == a) continue;
        const V r_ab = r[a] - r[b];
        const V gW = K::template grad<D>(r_ab, h);
        const double cs_ab = std::max(cs[a], cs[b]);
        const V Psi = (cs_ab * (rho[a] - rho[b]) * r_ab) / norm(r_ab);
        acc += m[b] / gamma[a] * dot((v[a] - v[b]) + Psi / rho[b], gW);
      }
      drho_dt[a] = acc;
    }
  }

  // ---- fluid_equations.hpp:267-305 ----
  void compute_momentum() {
    if (prm.symmetric) return compute_momentum_symmetric();
    const std::size_t nn = n();
    const double h = prm.h;
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nn; ++a) p[a] = eos_p(rho[a]);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      V acc{};
      acc[1] = -prm.g;
      for (std::uint32_t i = fc_off[a]; i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        const double rho_s = favg(rho, s), p_s = favg(p, s);
        const double P_as = rho_s * (p[a] / pow2(rho[a]) + p_s / pow2(rho_s));
        const V n_s = normalize(gg);
        const V v_as = v[a] - favg(v, s);
        const V t_as = normalize(v_as - dot(v_as, n_s) * n_s);
        const double dr_as = std::max(h / 2.0, dot(r[a] - favg(r, s), n_s));
        const V Pi_as = (2.0 * prm.mu / (rho[a] * dr_as) * dot(v_as, t_as)) * t_as;
        acc += (P_as * gg - Pi_as * norm(gg)) / gamma[a];
      }
      for (std::uint32_t i = nb_off[a]; i < nb_off[a + 1]; ++i) {
        const std::size_t b = nb[i];
        if (b == a) continue;
        const V r_ab = r[a] - r[b];
        const V gW = K::template grad<D>(r_ab, h);
        const double P_ab = p[a] / pow2(rho[a]) + p[b] / pow2(rho[b]);
        const double Pi_ab = 2.0 * prm.mu * dot(v[a] - v[b], r_ab) / (rho[a] * rho[b] * norm2(r_ab));
        acc += (m[b] / gamma[a] * (Pi_ab - P_ab)) * gW;
      }
      dv_dt[a] = acc;
    }
  }

  // ---- fluid_equations.hpp:331-471 ----
  void apply_shifts() {
    const std::size_t nn = n();
    const double h = prm.h;
    const double rad = radius();
    const bool sym = prm.symmetric != 0;
    if (sym) shift_sums_symmetric();
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nn; ++a) {
      V Na{}, gr{};
      M La{}, gv{};
      if (sym) { Na = dr[a]; La = L[a]; gv = grad_v[a]; gr = grad_rho[a]; }
      for (std::uint32_t i = fc_off[a]; !sym && i < fc_off[a + 1]; ++i) {
        const std::uint32_t s = fc[i];
        const V gg = K::flux(domain.face(s), r[a], h);
        Na -= gg / gamma[a];
        La -= outer(favg(r, s) - r[a], gg) / gamma[a];
        gv -= outer(favg(v, s) - v[a], gg) / gamma[a];
        gr -= ((favg(rho, s) - rho[a]) * gg) / gamma[a];
      }
      for (std::uint32_t i = nb_off[a]; !sym && i < nb_off[a + 1]; ++i) {
        const std::size_t b = nb[i];
        if (b == a) continue;
        const double V_b = m[b] / rho[b];
        const V gW = K::template grad<D>(r[a] - r[b], h);
        const double c = V_b / gamma[a];
        Na += c * gW;
        La += c * outer(r[b] - r[a], gW);
        gv += c * outer(v[b] - v[a], gW);
        gr += (c * (rho[b] - rho[a])) * gW;
      }
      // :366-377
      dr[a] = Na;
      M Linv;
      const bool lu_ok = lu_inverse(transpose(La), Linv);
      if (!lu_ok) {
#pragma omp atomic
        stats[0]++;
      }
      if (lu_ok) {
        La = Linv;
        Na = matvec(La, Na);
        gv = matmul(gv, transpose(La));
        gr = matvec(La, gr);
      } else {
        La = eye<D>();
      }
      N[a] = normalize(Na);
      L[a] = La;
      grad_v[a] = gv;
      grad_rho[a] = gr;
    }
    // :387-388
    for (std::size_t a = 0; a < nn; ++a) phi[a] = is_fluid(a) ? phi_min : phi_max;
    // :396-416 visibility, gather form (all neighbours are within the radius).
    const double cos_fov = std::cos(M_PI / 4);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      bool vis = false;
      for (std::uint32_t i = nb_off[a]; i < nb_off[a + 1] && !vis; ++i) {
        const std::size_t b = nb[i];
        if (b == a) continue;
        const V r_ab = r[a] - r[b];
        const double r2 = norm2(r_ab);
        if (r2 > pow2(rad)) continue;
        const double thr = pow2(cos_fov) * r2;
        const double n_a = dot(N[a], r_ab);
        if (n_a > 0 && pow2(n_a) >= thr) vis = true;
      }
      if (vis) phi[a] = phi_max;
    }
    // :419-426 splashes.
    const std::uint32_t cutoff = D == 2 ? 8 : 26;
    for (std::size_t a = 0; a < nf; ++a)
      if (nb_off[a + 1] - nb_off[a] <= cutoff) {
        if (!bitwise_equal(phi[a], phi_min)) stats[2]++;
        phi[a] = phi_min;
      }
    for (std::size_t a = 0; a < nf; ++a) stats[1] += bitwise_equal(phi[a], phi_min);
    // :440-452 near-surface scaling (two-phase: readers only compare with phi_min).
    std::vector<double> phi_new(phi);
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      if (!bitwise_equal(phi[a], phi_max)) continue;
      bool found = false;
      std::size_t best = 0;
      double best_d = 0.0;
      for (std::uint32_t i = nb_off[a]; i < nb_off[a + 1]; ++i) {
        const std::size_t b = nb[i];
        if (!bitwise_equal(phi[b], phi_min)) continue;
        const double d2 = norm2(r[a] - r[b]);
        if (!found || d2 < best_d) { found = true; best = b; best_d = d2; }
      }
      if (found) phi_new[a] = phi[a] * (std::abs(dot(N[best], r[a] - r[best])) / rad);
    }
    phi.swap(phi_new);
    for (std::size_t a = 0; a < nf; ++a) {
      stats[3] += !bitwise_equal(phi[a], phi_max) && !bitwise_equal(phi[a], phi_min);
      stats[4] += bitwise_equal(phi[a], phi_max);
      stats[5] += bitwise_equal(phi[a], phi_max) && !approx_equal(gamma[a], 1.0);
    }
    // :455-470
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nf; ++a) {
      if (!bitwise_equal(phi[a], phi_max)) { dr[a] = V{}; continue; }
      dr[a] = dr[a] * (-CFL * C_shift * pow2(h));
      r[a] += dr[a];
      if (approx_equal(gamma[a], 1.0)) v[a] += matvec(grad_v[a], dr[a]);
      rho[a] += dot(grad_rho[a], dr[a]);
    }
  }

  // ---- fluid_equations.hpp:484-512 ----
  void apply_free_surface_correction() {
    const std::size_t nn = n();
    const double h = prm.h;
    const double K_fs = -std::log(0.05) / pow2(0.01);
    for (std::size_t a = 0; a < nn; ++a) rho_raw[a] = rho[a];
#pragma omp parallel for schedule(dynamic, 256)
    for (std::size_t a = 0; a < nf; ++a) {
      if (bitwise_equal(phi[a], phi_max)) continue;
      double alpha = 0.0, rho_t = 0.0;
      for (std::uint32_t i = nb_off[a]; i < nb_off[a + 1]; ++i) {
        const std::size_t b = nb[i];
        const double W = K::template value<D>(r[a] - r[b], h);
        alpha += m[b] / rho_raw[b] * W;
        rho_t += m[b] * W;
      }
      const double ratio = std::min(1.0, alpha / gamma[a]);
#pragma omp atomic
      stats[7]++;
      if (ratio > 0.99) continue;
#pragma omp atomic
      stats[6]++;
      const double beta = std::exp(-K_fs * pow2(ratio - 1.0));
      const double corr = beta * gamma[a] + (1.0 - beta) * alpha;
      if (!is_tiny(corr)) rho[a] = rho_t / corr;
    }
  }

  void post_only() override { post_integrate(); }
  void post_integrate() {
    for (long long& x : stats) x = 0;
    prepare();
    { PhaseTimer t(phase_s[5]); apply_shifts(); }
    { PhaseTimer t(phase_s[6]); apply_free_surface_correction(); }
  }

  void rhs_only() override {
    prepare();
    compute_continuity();
    compute_momentum();
  }

  // ---- time_integrator.hpp ----
  double substep(double dt, bool have_dt) {
    prepare();
    if (!have_dt) { PhaseTimer t(phase_s[4]); dt = compute_time_step(); }
    { PhaseTimer t(phase_s[3]); compute_continuity(); compute_momentum(); }
    PhaseTimer t(phase_s[4]);
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nf; ++a) {
      r[a] += dt * v[a];
      v[a] += dt * dv_dt[a];
      rho[a] += dt * drho_dt[a];
    }
    return dt;
  }
  void lincomb(const std::vector<V>& r0, const std::vector<V>& v0, const std::vector<double>& rho0_, double w) {
    PhaseTimer t(phase_s[4]);
#pragma omp parallel for schedule(static)
    for (std::size_t a = 0; a < nf; ++a) {
      r[a] = (1 - w) * r0[a] + w * r[a];
      v[a] = (1 - w) * v0[a] + w * v[a];
      rho[a] = (1 - w) * rho0_[a] + w * rho[a];
    }
  }

  double step() override {
    double dt = 0.0;
    switch (prm.integrator) {
      case 0: {  // :49-69
        prepare();
        dt = compute_time_step();
        compute_continuity();
        for (std::size_t a = 0; a < nf; ++a) rho[a] += dt * drho_dt[a];
        compute_momentum();
        for (std::size_t a = 0; a < nf; ++a) { v[a] += dt * dv_dt[a]; r[a] += dt * v[a]; }
        post_integrate();
        break;
      }
      case 1: {  // :98-123
        prepare();
        dt = compute_time_step();
        const double dt2 = dt / 2;
        compute_momentum();
        for (std::size_t a = 0; a < nf; ++a) { v[a] += dt2 * dv_dt[a]; r[a] += dt * v[a]; }
        prepare();
        compute_continuity();
        for (std::size_t a = 0; a < nf; ++a) rho[a] += dt * drho_dt[a];
        compute_momentum();
        for (std::size_t a = 0; a < nf; ++a) v[a] += dt2 * dv_dt[a];
        post_integrate();
        break;
      }
      case 2:
      case 3: {  // :161-184
        const std::vector<V> r0(r), v0(v);
        const std::vector<double> rho_0(rho);
        dt = substep(0.0, false);
        if (prm.integrator == 2) {
          substep(dt, true);
          lincomb(r0, v0, rho_0, 1.0 / 2.0);
        } else {
          substep(dt, true);
          lincomb(r0, v0, rho_0, 1.0 / 4.0);
          substep(dt, true);
          lincomb(r0, v0, rho_0, 2.0 / 3.0);
        }
        post_integrate();
        break;
      }
      default: break;
    }
    return dt;
  }

  void neighbors(std::vector<std::uint64_t>& off, std::vector<std::uint64_t>& cols) override {
    search();
    off.assign(nb_off.begin(), nb_off.end());
    cols.assign(nb.begin(), nb.end());
  }
  void face_neighbors(std::vector<std::uint64_t>& off, std::vector<std::uint64_t>& cols) override {
    search();
    off.assign(fc_off.begin(), fc_off.end());
    cols.assign(fc.begin(), fc.end());
  }
};

}  // namespace orc
