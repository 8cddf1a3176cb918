// tit_b200/sph.hpp — C++ facade over the titgpu C ABI (include/titgpu.h).
//
// Keeps the names, call order and argument meaning of the reference's
// header-only template API for the WCSPH particle step, so that
// /root/reference/source/titwcsph/wcsph.cpp compiles nearly verbatim against
// the B200 library (see examples/dam_break_2d.cpp and INTEGRATION.md):
//
//   tit::Vec, tit::Mat                              tit/core/vec.hpp, mat.hpp (the subset the driver uses)
//   tit::geom::Surface, tessellate (2-D, 3-D)       tit/geom/surface.hpp, tessellation.hpp:30-63, 74-163
//   tit::geom::MakeFastWinding                      tit/geom/winding/fast_winding.hpp (exact winding, exact_winding.hpp:32-43)
//   tit::geom::GridSearch, GridFaceSearch, ...      tit/geom/search.hpp, face_search.hpp, partition.hpp (option holders)
//   tit::sph::Space, ParticleType, field tags       tit/sph/field.hpp:112-218
//   tit::sph::ParticleArray / ParticleView          tit/sph/particle_array.hpp:41-293
//   tit::sph::ParticleMesh                          tit/sph/particle_mesh.hpp:42-258
//   tit::sph::*Kernel, Tait / LinearTait EOS        tit/sph/kernel.hpp:428-454, equation_of_state.hpp:19-122
//   tit::sph::FluidEquations                        tit/sph/fluid_equations.hpp:37-533
//   tit::sph::SSPRKIntegrator, SymplecticEuler..., VelocityVerlet...   tit/sph/time_integrator.hpp:32-228
//
// The physics runs on the GPU; this header only owns host mirrors of the
// particle fields and synchronises them lazily (upload before a step if the
// host copy was written, download on first access after a step). Errors of the
// C ABI are rethrown as tit::Exception (the reference's TIT_ENSURE behaviour,
// tit/core/exception.hpp:77-84). C++20, header-only, no CUDA headers needed.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <ranges>
#include <span>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../titgpu.h"
#include "core.hpp"
#include "data.hpp"

namespace tit {

namespace par {
/// tit/par/control.hpp: the CPU thread pool has no GPU counterpart.
inline void init() noexcept {}
}  // namespace par

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
namespace geom {

/// tit/geom/surface.hpp:28-111 (vertices + faces of Dim vertex indices).
template<class V>
class Surface final {
public:
  static constexpr auto Dim = vec_dim_v<V>;
  using FaceVerts = std::array<std::size_t, Dim>;
  auto num_verts() const noexcept -> std::size_t { return verts_.size(); }
  auto vert(std::size_t i) const noexcept -> const V& { return verts_[i]; }
  auto verts() const noexcept -> std::span<const V> { return verts_; }
  void append_vert(const V& v) { verts_.push_back(v); }
  auto num_faces() const noexcept -> std::size_t { return faces_.size(); }
  auto face_verts(std::size_t f) const noexcept -> const FaceVerts& { return faces_[f]; }
  auto face_verts() const noexcept -> std::span<const FaceVerts> { return faces_; }
  /// The face as geometry: its vertices in order (the reference returns a Segment / Triangle built from them, tit/geom/surface.hpp).
  auto face(std::size_t f) const noexcept -> std::array<V, Dim> {
    std::array<V, Dim> out;
    for (std::size_t k = 0; k < Dim; ++k) out[k] = verts_[faces_[f][k]];
    return out;
  }
  void append_face(const FaceVerts& f) { faces_.push_back(f); }
private:
  std::vector<V> verts_;
  std::vector<FaceVerts> faces_;
};

/// tit/geom/tessellation.hpp:30-63: split every segment into ceil(len / d_max) parts.
template<class V>
  requires (vec_dim_v<V> == 2)
auto tessellate(const Surface<V>& surf, vec_num_t<V> d_max) -> Surface<V> {
  using Num = vec_num_t<V>;
  Surface<V> result;
  for (const auto& vert : surf.verts()) result.append_vert(vert);
  for (std::size_t f = 0; f < surf.num_faces(); ++f) {
    auto [prev, last] = surf.face_verts(f);
    const V a = surf.vert(prev), ba = surf.vert(last) - a;
    const Num d = std::sqrt(ba[1] * ba[1] + ba[0] * ba[0]);  // norm(cross(ba)), geom/segment.hpp:71-73
    const auto n = std::max<std::size_t>(1, static_cast<std::size_t>(std::ceil(d / d_max)));
    for (std::size_t i = 1; i < n; ++i) {
      const Num t = static_cast<Num>(i) / static_cast<Num>(n);
      result.append_vert(a + t * ba);
      const auto vi = result.num_verts() - 1;
      result.append_face({prev, vi});
      prev = vi;
    }
    result.append_face({prev, last});
  }
  return result;
}

/// tit/geom/tessellation.hpp:74-163: red refinement of a triangle surface until no
/// edge is longer than d_max. Every sweep gives each too-long edge one midpoint
/// (appended in the order the edges are met: faces in order, edges ab, bc, ca) and
/// cuts every triangle along its midpoints; vertex and face numbering are those of
/// the reference (geom/tessellation.test.cpp:64-215). Host-side set-up code.
template<class V>
  requires (vec_dim_v<V> == 3)
auto tessellate(const Surface<V>& surf, vec_num_t<V> d_max) -> Surface<V> {
  using Num = vec_num_t<V>;
  using Tri = std::array<std::size_t, 3>;
  constexpr auto none = static_cast<std::size_t>(-1);
  std::vector<V> verts(surf.verts().begin(), surf.verts().end());
  std::vector<Tri> faces(surf.face_verts().begin(), surf.face_verts().end()), next;
  std::map<std::pair<std::size_t, std::size_t>, std::size_t> mid;
  const auto key = [](std::size_t i, std::size_t j) { return i < j ? std::pair{i, j} : std::pair{j, i}; };
  for (;;) {
    mid.clear();
    for (const Tri& t : faces) {
      for (int k = 0; k < 3; ++k) {
        const std::size_t i = t[k], j = t[(k + 1) % 3];
        if (mid.contains(key(i, j))) continue;
        const V e = verts[j] - verts[i];
        if (dot(e, e) <= d_max * d_max) continue;
        verts.push_back((verts[i] + verts[j]) / Num{2});
        mid.emplace(key(i, j), verts.size() - 1);
      }
    }
    if (mid.empty()) break;
    next.clear();
    for (const Tri& t : faces) {
      std::size_t m[3];
      int n_split = 0;
      for (int k = 0; k < 3; ++k) {
        const auto it = mid.find(key(t[k], t[(k + 1) % 3]));
        m[k] = it == mid.end() ? none : it->second;
        n_split += m[k] != none;
      }
      if (n_split == 0) { next.push_back(t); continue; }
      // Rotate so that (p, q) is the first split edge in cyclic order.
      int r = 0;
      if (n_split < 3) while (!(m[r] != none && m[(r + 2) % 3] == none)) ++r;
      const std::size_t p = t[r], q = t[(r + 1) % 3], s = t[(r + 2) % 3];
      const std::size_t m_pq = m[r], m_qs = m[(r + 1) % 3], m_sp = m[(r + 2) % 3];
      if (n_split == 3) {
        next.push_back({p, m_pq, m_sp}); next.push_back({m_pq, q, m_qs}); next.push_back({m_sp, m_qs, s}); next.push_back({m_pq, m_qs, m_sp});
      } else if (n_split == 2) {
        next.push_back({p, m_pq, s}); next.push_back({m_pq, m_qs, s}); next.push_back({m_pq, q, m_qs});
      } else {
        next.push_back({p, m_pq, s}); next.push_back({m_pq, q, s});
      }
    }
    faces.swap(next);
  }
  Surface<V> result;
  for (const V& v : verts) result.append_vert(v);
  for (const Tri& f : faces) result.append_face(f);
  return result;
}

/// Containment functor (tit/geom/winding/fast_winding.hpp:52-92). The tree of
/// the reference is an accelerator that falls back to the exact generalized
/// winding number; the GPU evaluates the exact number where a cell is cut by
/// the surface, so only the exact form is kept. Keeps a pointer to the
/// surface: temporaries are rejected as in the reference (:371).
template<class V>
class WindingFunc final {
public:
  explicit WindingFunc(const Surface<V>& surf) noexcept : surf_{&surf} {}
  explicit WindingFunc(Surface<V>&&) = delete;
  auto surface() const noexcept -> const Surface<V>& { return *surf_; }
  auto operator()(const V& p) const noexcept -> vec_num_t<V> {
    using Num = vec_num_t<V>;
    Num w{};
    for (const auto& fv : surf_->face_verts()) {
      if constexpr (vec_dim_v<V> == 2) {  // geom/segment.hpp:76-81
        const V ap = surf_->vert(fv[0]) - p, bp = surf_->vert(fv[1]) - p;
        w += std::atan2(ap[0] * bp[1] - ap[1] * bp[0], dot(ap, bp)) / (2 * M_PI);
      } else {  // geom/triangle.hpp:93-103
        const V ap = surf_->vert(fv[0]) - p, bp = surf_->vert(fv[1]) - p, cp = surf_->vert(fv[2]) - p;
        const Num an = norm(ap), bn = norm(bp), cn = norm(cp);
        const Num den = an * bn * cn + dot(ap, bp) * cn + dot(bp, cp) * an + dot(cp, ap) * bn;
        const Num det = ap[0] * (bp[1] * cp[2] - bp[2] * cp[1]) + ap[1] * (bp[2] * cp[0] - bp[0] * cp[2]) + ap[2] * (bp[0] * cp[1] - bp[1] * cp[0]);
        w += std::atan2(det, den) / (2 * M_PI);
      }
    }
    return w;
  }
  auto contains(const V& p) const noexcept -> bool { return (*this)(p) > 0.5; }
private:
  const Surface<V>* surf_;
};
template<class Num>
struct MakeFastWinding final {
  template<class V> auto operator()(const Surface<V>& surf) const { return WindingFunc<V>{surf}; }
  template<class V> void operator()(Surface<V>&&) const = delete;
};
template<class V> auto make_exact_winding(const Surface<V>& surf) { return WindingFunc<V>{surf}; }

/// Search / partition option holders (tit/geom/search/grid_search.hpp:117-140,
/// face_search/grid_face_search.hpp:158-181, partition/*.hpp). The cell hints
/// are passed on for interface parity; the GPU hash chooses its own cell size
/// (neighbour sets do not depend on it) and needs no block partition.
template<class Num> struct GridSearch final { Num size_hint; };
template<class Num> GridSearch(Num) -> GridSearch<Num>;
template<class Num> struct GridFaceSearch final { Num size_hint; };
template<class Num> GridFaceSearch(Num) -> GridFaceSearch<Num>;
/// K-d tree index option (search/kd_tree_search.hpp:129-142): same contract as the grid
/// index; on the GPU either is served by the spatial hash (identical neighbour rows,
/// tests/test_gpu_parity.py::test_neighbors_match_kd_tree_index).
struct KDTreeSearch final {};
inline constexpr KDTreeSearch kd_tree_indexing{};
/// Partitioners (partition/recursive_bisection.hpp:102-113, sort_partition.hpp:82-93,
/// kmeans_clustering.hpp:30-138): they colour pairs for the CPU thread pool; the
/// gather-form GPU pair sums need no colouring, so the options are accepted and unused.
struct RecursiveInertialBisection final {};
struct RecursiveCoordBisection final {};
struct HilbertCurvePartition final {};
struct MortonCurvePartition final {};
struct KMeansClustering final {
  constexpr explicit KMeansClustering(float64_t eps = 1.0e-4, std::size_t max_iter = 10) noexcept : eps{eps}, max_iter{max_iter} {}
  float64_t eps;
  std::size_t max_iter;
};
inline constexpr RecursiveInertialBisection recursive_inertial_bisection{};
inline constexpr RecursiveCoordBisection recursive_coord_bisection{};
inline constexpr HilbertCurvePartition hilbert_curve_partition{};
inline constexpr MortonCurvePartition morton_curve_partition{};
inline constexpr KMeansClustering kmeans_clustering{};
template<class Num, class Clustering = KMeansClustering> struct PixelatedPartition final { Num size_hint; Clustering clustering{}; };
template<class Num, class C> PixelatedPartition(Num, C) -> PixelatedPartition<Num, C>;

}  // namespace geom

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Fields (tit/sph/field.hpp:112-218). `r` and `phi` live in tit::sph, the rest in tit.

namespace sph {
template<class Num, std::size_t Dim> struct Space {};
enum class ParticleType : std::uint8_t { fluid, fixed, count };
template<class Real, std::size_t Dim> class ParticleArray;
template<class Array> class ParticleView;
}  // namespace sph

namespace impl {
enum class Rank { scalar, vector, matrix };
/// Field tag. `id` is the position in the reference's varying-field list
/// (fluid_equations.hpp:41-48); -1 = the uniform field h.
template<int Id, Rank R>
struct Field {
  static constexpr int id = Id;
  static constexpr Rank rank = R;
  std::string_view field_name;
  template<class PV> constexpr auto operator[](PV&& a) const noexcept -> decltype(auto) { return std::forward<PV>(a)[*this]; }
  /// f[a, b] = f[a] - f[b] (field.hpp:51-54); `f(a, b)` where multi-argument subscripts (C++23) are unavailable.
  template<class PVa, class PVb> constexpr auto operator()(PVa&& a, PVb&& b) const noexcept { return a[*this] - b[*this]; }
#if defined(__cpp_multidimensional_subscript)
  template<class PVa, class PVb> constexpr auto operator[](PVa&& a, PVb&& b) const noexcept { return a[*this] - b[*this]; }
#endif
};
}  // namespace impl

#define TIT_B200_FIELD(name, id, rank) inline constexpr impl::Field<id, impl::Rank::rank> name{#name}
TIT_B200_FIELD(h, -1, scalar);
TIT_B200_FIELD(m, 0, scalar);
TIT_B200_FIELD(gamma, 1, scalar);
TIT_B200_FIELD(grad_gamma, 2, vector);
TIT_B200_FIELD(rho, 3, scalar);
TIT_B200_FIELD(drho_dt, 4, scalar);
TIT_B200_FIELD(grad_rho, 5, vector);
TIT_B200_FIELD(p, 6, scalar);
TIT_B200_FIELD(cs, 7, scalar);
TIT_B200_FIELD(v, 8, vector);
TIT_B200_FIELD(dv_dt, 9, vector);
TIT_B200_FIELD(grad_v, 10, matrix);
namespace sph { TIT_B200_FIELD(r, 11, vector); }
TIT_B200_FIELD(dr, 12, vector);
TIT_B200_FIELD(L, 13, matrix);
TIT_B200_FIELD(N, 14, vector);
namespace sph { TIT_B200_FIELD(phi, 15, scalar); }
TIT_B200_FIELD(rho_raw, 16, scalar);
#undef TIT_B200_FIELD

namespace sph {

inline constexpr int num_varying_fields = 17;
inline constexpr std::array<const char*, num_varying_fields> varying_field_names{
    "m", "gamma", "grad_gamma", "rho", "drho_dt", "grad_rho", "p", "cs", "v", "dv_dt", "grad_v", "r", "dr", "L", "N", "phi", "rho_raw"};
inline constexpr std::array<impl::Rank, num_varying_fields> varying_field_ranks{
    impl::Rank::scalar, impl::Rank::scalar, impl::Rank::vector, impl::Rank::scalar, impl::Rank::scalar, impl::Rank::vector, impl::Rank::scalar, impl::Rank::scalar, impl::Rank::vector,
    impl::Rank::vector, impl::Rank::matrix, impl::Rank::vector, impl::Rank::vector, impl::Rank::matrix, impl::Rank::vector, impl::Rank::scalar, impl::Rank::scalar};

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Kernels and equations of state: tag types carrying the C ABI ids.

struct CubicSplineKernel final { static constexpr int id = TITGPU_KERNEL_CUBIC_SPLINE; };
struct QuarticSplineKernel final { static constexpr int id = TITGPU_KERNEL_QUARTIC_SPLINE; };
struct QuinticSplineKernel final { static constexpr int id = TITGPU_KERNEL_QUINTIC_SPLINE; };
struct QuarticWendlandKernel final { static constexpr int id = TITGPU_KERNEL_QUARTIC_WENDLAND; };
struct SixthOrderWendlandKernel final { static constexpr int id = TITGPU_KERNEL_SIXTH_ORDER_WENDLAND; };
struct EighthOrderWendlandKernel final { static constexpr int id = TITGPU_KERNEL_EIGHTH_ORDER_WENDLAND; };

/// equation_of_state.hpp:19-72.
template<class Num>
class TaitEquationOfState final {
public:
  static constexpr int id = TITGPU_EOS_TAIT;
  constexpr explicit TaitEquationOfState(Num cs_0, Num rho_0, Num xi = Num{7}) noexcept : cs_0_{cs_0}, rho_0_{rho_0}, xi_{xi} {}
  constexpr auto cs_0() const noexcept { return cs_0_; }
  constexpr auto rho_0() const noexcept { return rho_0_; }
  constexpr auto xi() const noexcept { return xi_; }
private:
  Num cs_0_, rho_0_, xi_;
};
/// equation_of_state.hpp:78-122.
template<class Num>
class LinearTaitEquationOfState final {
public:
  static constexpr int id = TITGPU_EOS_LINEAR_TAIT;
  constexpr explicit LinearTaitEquationOfState(Num cs_0, Num rho_0) noexcept : cs_0_{cs_0}, rho_0_{rho_0} {}
  constexpr auto cs_0() const noexcept { return cs_0_; }
  constexpr auto rho_0() const noexcept { return rho_0_; }
  constexpr auto xi() const noexcept { return Num{1}; }
private:
  Num cs_0_, rho_0_;
};

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
/// FluidEquations (fluid_equations.hpp:37-533). Stores the surfaces by value
/// (:528-531); `initialize` binds them to the particle array's GPU context.
template<class Real, std::size_t Dim, class EOS, class Kernel>
class FluidEquations final {
public:
  using V = Vec<Real, Dim>;
  static constexpr std::size_t dim = Dim;
  static constexpr int kernel_id = Kernel::id;
  static constexpr int eos_id = EOS::id;
  FluidEquations(Real g, Real mu, const geom::Surface<V>& domain, const geom::WindingFunc<V>& containment, EOS eos, Kernel /*kernel*/)
      : g_{g}, mu_{mu}, domain_{domain}, containment_{containment.surface()}, eos_{eos} {}
  auto g() const noexcept { return g_; }
  auto mu() const noexcept { return mu_; }
  auto eos() const noexcept -> const EOS& { return eos_; }
  auto domain() const noexcept -> const geom::Surface<V>& { return domain_; }
  auto containment() const noexcept -> const geom::Surface<V>& { return containment_; }

  /// fluid_equations.hpp:79-89.
  template<class Mesh> void initialize(Mesh& mesh, ParticleArray<Real, Dim>& particles) const {
    particles.bind_(*this, mesh);
    particles.push_();
    particles.call_(titgpu_initialize(particles.ctx_()), "titgpu_initialize");
    particles.mark_device_newer_();
  }
  /// fluid_equations.hpp:99-105 (neighbour search + gamma + wall particles).
  template<class Mesh> void prepare(Mesh& mesh, ParticleArray<Real, Dim>& particles) const {
    particles.bind_(*this, mesh);
    particles.push_();
    particles.call_(titgpu_prepare(particles.ctx_()), "titgpu_prepare");
    particles.mark_device_newer_();
  }
private:
  Real g_, mu_;
  geom::Surface<V> domain_, containment_;
  EOS eos_;
};
template<class Real, std::size_t Dim, class EOS, class Kernel>
FluidEquations(Real, Real, const geom::Surface<Vec<Real, Dim>>&, const geom::WindingFunc<Vec<Real, Dim>>&, EOS, Kernel) -> FluidEquations<Real, Dim, EOS, Kernel>;

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Time integrators (time_integrator.hpp:32-228).

enum class SSPRKOrder : std::uint8_t { two = 2, three = 3 };

namespace impl {
template<class Equations>
class IntegratorBase {
public:
  using equations_type = Equations;
  auto equations() const noexcept -> const Equations& { return equations_; }
  auto integrator_id() const noexcept -> int { return id_; }
  /// One time step; returns dt. Mutates the mesh and the particles, as in the reference.
  template<class Mesh, class Real, std::size_t Dim>
  auto step(Mesh& mesh, ParticleArray<Real, Dim>& particles) const -> Real {
    particles.require_integrator_(id_);
    particles.bind_(equations_, mesh);
    particles.push_();
    double dt = 0.0;
    particles.call_(titgpu_step(particles.ctx_(), 1, &dt), "titgpu_step");
    particles.mark_device_newer_();
    mesh.invalidate_();
    return static_cast<Real>(dt);
  }
protected:
  IntegratorBase(Equations equations, int id) : equations_{std::move(equations)}, id_{id} {}
private:
  Equations equations_;
  int id_;
};
}  // namespace impl

template<class Equations>
class SymplecticEulerIntegrator final : public impl::IntegratorBase<Equations> {
public:
  explicit SymplecticEulerIntegrator(Equations equations) : impl::IntegratorBase<Equations>{std::move(equations), TITGPU_SYMPLECTIC_EULER} {}
};
template<class Equations>
class VelocityVerletIntegrator final : public impl::IntegratorBase<Equations> {
public:
  explicit VelocityVerletIntegrator(Equations equations) : impl::IntegratorBase<Equations>{std::move(equations), TITGPU_VELOCITY_VERLET} {}
};
template<class Equations>
class SSPRKIntegrator final : public impl::IntegratorBase<Equations> {
public:
  explicit SSPRKIntegrator(Equations equations, SSPRKOrder order = SSPRKOrder::three)
      : impl::IntegratorBase<Equations>{std::move(equations), order == SSPRKOrder::two ? TITGPU_SSPRK2 : TITGPU_SSPRK3} {}
};

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
/// ParticleView (particle_array.hpp:41-124).
template<class Array>
class ParticleView final {
public:
  constexpr ParticleView(Array& array, std::size_t index) noexcept : array_{&array}, index_{index} {}
  constexpr auto array() const noexcept -> Array& { return *array_; }
  constexpr auto index() const noexcept -> std::size_t { return index_; }
  constexpr auto has_type(ParticleType type) const noexcept -> bool { return array_->has_type(index_, type); }
  constexpr auto is_fluid() const noexcept -> bool { return has_type(ParticleType::fluid); }
  constexpr auto is_fixed() const noexcept -> bool { return has_type(ParticleType::fixed); }
  template<int Id, tit::impl::Rank R>
  auto operator[](tit::impl::Field<Id, R> field) const -> decltype(auto) { return array_->at_(index_, field); }
  friend constexpr auto operator==(ParticleView a, ParticleView b) noexcept -> bool { return a.index_ == b.index_; }
private:
  Array* array_;
  std::size_t index_;
};

/// What a step publishes for read-back (titgpu_set_outputs).
enum class Publish : int { state = 0, fluid = 1, all = 2 };

/// ParticleArray (particle_array.hpp:139-293): host SoA mirror + the GPU context.
/// Particles are ordered by type: fluid first, then fixed (:188-199, :233-239).
template<class Real, std::size_t Dim>
class ParticleArray final {
  static_assert(std::is_same_v<Real, double>, "the B200 path computes in fp64 (wcsph.cpp:205)");
public:
  using V = Vec<Real, Dim>;
  using M = Mat<Real, Dim>;

  /// `ParticleArray particles{Space<Real, Dim>{}, time_integrator};` (wcsph.cpp:96-101).
  template<class Integrator>
  ParticleArray(Space<Real, Dim> /*space*/, const Integrator& integrator, int device = 0)
      : kernel_id_{Integrator::equations_type::kernel_id}, eos_id_{Integrator::equations_type::eos_id}, integrator_id_{integrator.integrator_id()}, device_{device} {
    static_assert(Integrator::equations_type::dim == Dim);
  }
  ParticleArray(const ParticleArray&) = delete;
  auto operator=(const ParticleArray&) -> ParticleArray& = delete;
  ~ParticleArray() { if (ctx_raw_ != nullptr) titgpu_destroy(ctx_raw_); }

  auto size() const noexcept -> std::size_t { return ranges_[2]; }
  void reserve(std::size_t capacity) { for (int f = 0; f < num_varying_fields; ++f) cols_[f].reserve(capacity * width_(f)); }

  /// Write all varying fields as one frame of `series` (particle_array.hpp:165-172):
  /// one array per field, named after the field, in the field-set order. Needs
  /// `publish(Publish::all)` during the step before (the default).
  void write(Real time, data::SeriesView<data::Storage> series) const {
    if (publish_ != Publish::all && stepped_) throw Exception("ParticleArray::write: the last step did not publish every field (publish(Publish::all) before it)");
    const auto frame = series.create_frame(static_cast<float64_t>(time));
    for (int f = 0; f < num_varying_fields; ++f) {
      fetch_(f);
      const auto rank = static_cast<data::Rank>(varying_field_ranks[f]);
      const data::Type type{data::kind_of<Real>, rank, static_cast<std::uint8_t>(rank == data::Rank::scalar ? 1 : Dim)};
      frame.create_array(varying_field_names[f]).write(type, std::as_bytes(std::span<const Real>{cols_[f]}));
    }
  }

  /// particle_array.hpp:188-199.
  auto append(ParticleType type) -> ParticleView<ParticleArray> {
    pull_all_();
    const auto t = static_cast<std::size_t>(type);
    const std::size_t index = ranges_[t + 1];
    for (std::size_t k = t + 1; k < ranges_.size(); ++k) ranges_[k] += 1;
    for (int f = 0; f < num_varying_fields; ++f) cols_[f].insert(cols_[f].begin() + std::ptrdiff_t(index * width_(f)), width_(f), Real{});
    host_dirty_.fill(true);
    resized_ = true;
    return ParticleView<ParticleArray>{*this, index};
  }
  auto has_type(std::size_t index, ParticleType type) const noexcept -> bool {
    const auto t = static_cast<std::size_t>(type);
    return ranges_[t] <= index && index < ranges_[t + 1];
  }
  auto operator[](std::size_t index) noexcept { return ParticleView<ParticleArray>{*this, index}; }
  auto typed(ParticleType type) noexcept {
    const auto t = static_cast<std::size_t>(type);
    return std::views::iota(ranges_[t], ranges_[t + 1]) | std::views::transform([this](std::size_t i) { return ParticleView<ParticleArray>{*this, i}; });
  }
  auto all() noexcept { return std::views::iota(std::size_t{0}, size()) | std::views::transform([this](std::size_t i) { return ParticleView<ParticleArray>{*this, i}; }); }
  auto fluid() noexcept { return typed(ParticleType::fluid); }
  auto fixed() noexcept { return typed(ParticleType::fixed); }
  auto num_fluid() const noexcept -> std::size_t { return ranges_[1]; }
  auto num_fixed() const noexcept -> std::size_t { return ranges_[2] - ranges_[1]; }

  /// `field[particles]`: the uniform value (h) or a span over all particles (:264-276).
  template<int Id, tit::impl::Rank R>
  auto operator[](tit::impl::Field<Id, R> /*field*/) -> decltype(auto) {
    if constexpr (Id < 0) { params_dirty_ = true; return (h_); }
    else {
      touch_(Id);
      if constexpr (R == tit::impl::Rank::scalar) return std::span<Real>{cols_[Id]};
      else if constexpr (R == tit::impl::Rank::vector) return std::span<V>{reinterpret_cast<V*>(cols_[Id].data()), size()};
      else return std::span<M>{reinterpret_cast<M*>(cols_[Id].data()), size()};
    }
  }

  /// Read-only `field[particles]` (const array): fetches the device copy if it is newer and does NOT
  /// mark the host copy as written - `std::as_const(particles)[r]` between steps costs no re-upload.
  template<int Id, tit::impl::Rank R>
  auto operator[](tit::impl::Field<Id, R> /*field*/) const -> decltype(auto) {
    if constexpr (Id < 0) { return static_cast<const Real&>(h_); }
    else {
      fetch_(Id);
      if constexpr (R == tit::impl::Rank::scalar) return std::span<const Real>{cols_[Id]};
      else if constexpr (R == tit::impl::Rank::vector) return std::span<const V>{reinterpret_cast<const V*>(cols_[Id].data()), size()};
      else return std::span<const M>{reinterpret_cast<const M*>(cols_[Id].data()), size()};
    }
  }

  /// Which derived fields a step publishes (default: everything, as the reference).
  void publish(Publish level) { publish_ = level; if (ctx_raw_ != nullptr) call_(titgpu_set_outputs(ctx_raw_, static_cast<int>(level)), "titgpu_set_outputs"); }

  /// Sorted adjacency of the last search as CSR over original indices (ParticleMesh uses it).
  void neighbors_(std::vector<std::uint64_t>& offsets, std::vector<std::uint64_t>& cols) {
    push_();
    std::size_t nnz = 0;
    call_(titgpu_neighbors(ctx_(), nullptr, nullptr, 0, &nnz), "titgpu_neighbors");
    offsets.assign(size() + 1, 0);
    cols.assign(std::max<std::size_t>(nnz, 1), 0);
    call_(titgpu_neighbors(ctx_(), offsets.data(), cols.data(), nnz, &nnz), "titgpu_neighbors");
    cols.resize(nnz);
  }

  /// Sorted face adjacency of the last search (`mesh[domain, a]`), CSR over original particle indices.
  void face_neighbors_(std::vector<std::uint64_t>& offsets, std::vector<std::uint64_t>& cols) {
    push_();
    std::size_t nnz = 0;
    call_(titgpu_face_neighbors(ctx_(), nullptr, nullptr, 0, &nnz), "titgpu_face_neighbors");
    offsets.assign(size() + 1, 0);
    cols.assign(std::max<std::size_t>(nnz, 1), 0);
    call_(titgpu_face_neighbors(ctx_(), offsets.data(), cols.data(), nnz, &nnz), "titgpu_face_neighbors");
    cols.resize(nnz);
  }
  /// Fixed particle k (= vertex k of the domain surface, wcsph.cpp:110-113).
  auto fixed_at_(std::size_t k) noexcept { return ParticleView<ParticleArray>{*this, ranges_[1] + k}; }

  // ---- used by the equations / integrators (not part of the reference surface) ----
  auto ctx_() -> titgpu_ctx* {
    if (ctx_raw_ == nullptr) {
      const int rc = titgpu_create(&ctx_raw_, device_, int(Dim), kernel_id_, eos_id_, integrator_id_);
      if (rc != 0) {
        const std::string msg = ctx_raw_ != nullptr ? titgpu_last_error(ctx_raw_) : "out of memory";
        if (ctx_raw_ != nullptr) { titgpu_destroy(ctx_raw_); ctx_raw_ = nullptr; }
        throw Exception("titgpu_create: " + msg);
      }
      call_(titgpu_set_outputs(ctx_raw_, static_cast<int>(publish_)), "titgpu_set_outputs");
    }
    return ctx_raw_;
  }
  void call_(int rc, const char* what) const {
    if (rc != 0) throw Exception(std::string(what) + ": " + titgpu_last_error(ctx_raw_));
  }
  void require_integrator_(int id) {
    if (id != integrator_id_) throw Exception("the particle array was built for another time integrator");
  }
  template<class Equations, class Mesh>
  void bind_(const Equations& eq, const Mesh& mesh) {
    // (a mere READ of h through the non-const accessor sets params_dirty_: rebind - which rebuilds the
    // search grid and the wall cache - only if the value really changed)
    if (bound_ == static_cast<const void*>(&eq) && (!params_dirty_ || h_ == bound_h_)) { params_dirty_ = false; return; }
    using Eq = std::remove_cvref_t<Equations>;
    static_assert(Eq::dim == Dim);
    if (Eq::kernel_id != kernel_id_ || Eq::eos_id != eos_id_) throw Exception("the particle array was built for other equations");
    call_(titgpu_set_params(ctx_(), eq.g(), eq.mu(), eq.eos().cs_0(), eq.eos().rho_0(), eq.eos().xi(), h_, mesh.search_hint(), mesh.face_search_hint()), "titgpu_set_params");
    const auto flat = [](const geom::Surface<V>& s, std::vector<double>& verts, std::vector<std::uint64_t>& faces) {
      verts.clear(); faces.clear();
      for (const auto& q : s.verts()) for (std::size_t d = 0; d < Dim; ++d) verts.push_back(q[d]);
      for (const auto& f : s.face_verts()) for (std::size_t d = 0; d < Dim; ++d) faces.push_back(f[d]);
    };
    std::vector<double> dv, cv;
    std::vector<std::uint64_t> df, cf;
    flat(eq.domain(), dv, df);
    flat(eq.containment(), cv, cf);
    call_(titgpu_set_surface(ctx_(), dv.data(), eq.domain().num_verts(), df.data(), eq.domain().num_faces(), cv.data(), eq.containment().num_verts(), cf.data(),
                             eq.containment().num_faces()),
          "titgpu_set_surface");
    bound_ = &eq;
    bound_h_ = h_;
    params_dirty_ = false;
  }
  /// Upload the host-written input fields (state + the dv_dt seed of the time-step limit).
  void push_() {
    constexpr int inputs[] = {11 /*r*/, 8 /*v*/, 3 /*rho*/, 0 /*m*/, 9 /*dv_dt*/};
    for (const int f : inputs) {
      if (!host_dirty_[f] && !resized_) continue;
      if (f == 9 && !host_dirty_[f]) continue;
      call_(titgpu_upload(ctx_(), num_fluid(), num_fixed(), varying_field_names[f], cols_[f].data(), 0), "titgpu_upload");
      host_dirty_[f] = false;
    }
    resized_ = false;
  }
  void mark_device_newer_() { device_newer_.fill(true); stepped_ = true; }

  template<int Id, tit::impl::Rank R>
  auto at_(std::size_t index, tit::impl::Field<Id, R> /*field*/) -> decltype(auto) {
    if constexpr (Id < 0) { params_dirty_ = true; return (h_); }
    else {
      touch_(Id);
      if constexpr (R == tit::impl::Rank::scalar) return (cols_[Id][index]);
      else if constexpr (R == tit::impl::Rank::vector) return (reinterpret_cast<V*>(cols_[Id].data())[index]);
      else return (reinterpret_cast<M*>(cols_[Id].data())[index]);
    }
  }

private:
  static constexpr auto width_(int f) noexcept -> std::size_t {
    return varying_field_ranks[f] == tit::impl::Rank::scalar ? 1 : varying_field_ranks[f] == tit::impl::Rank::vector ? Dim : Dim * Dim;
  }
  /// Host access to field f: fetch the device copy if it is newer, then assume a write.
  void touch_(int f) {
    if (device_newer_[f]) {
      device_newer_[f] = false;
      if (ctx_raw_ != nullptr && size() > 0) call_(titgpu_download(ctx_raw_, varying_field_names[f], cols_[f].data(), 0), "titgpu_download");
    }
    host_dirty_[f] = true;
  }
  /// Read-only host access: fetch the device copy if it is newer.
  void fetch_(int f) const {
    if (!device_newer_[f]) return;
    device_newer_[f] = false;
    if (ctx_raw_ != nullptr && size() > 0) call_(titgpu_download(ctx_raw_, varying_field_names[f], cols_[f].data(), 0), "titgpu_download");
  }
  void pull_all_() { for (int f = 0; f < num_varying_fields; ++f) if (device_newer_[f]) { touch_(f); } }

  int kernel_id_, eos_id_, integrator_id_, device_;
  titgpu_ctx* ctx_raw_ = nullptr;
  const void* bound_ = nullptr;
  Real bound_h_{};
  mutable bool params_dirty_ = true;
  bool resized_ = true;
  Publish publish_ = Publish::all;
  Real h_{};
  std::array<std::size_t, 3> ranges_{0, 0, 0};
  mutable std::array<std::vector<Real>, num_varying_fields> cols_;  // (mutable: a const read may fetch the device copy)
  std::array<bool, num_varying_fields> host_dirty_{};
  mutable std::array<bool, num_varying_fields> device_newer_{};
  bool stepped_ = false;
};
template<class Real, std::size_t Dim, class Integrator>
ParticleArray(Space<Real, Dim>, const Integrator&) -> ParticleArray<Real, Dim>;
template<class Real, std::size_t Dim, class Integrator>
ParticleArray(Space<Real, Dim>, const Integrator&, int) -> ParticleArray<Real, Dim>;

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
/// ParticleMesh (particle_mesh.hpp:42-258): adjacency, fetched from the GPU on demand.
template<class Search, class FaceSearch, class Partition = geom::RecursiveInertialBisection, class InterfacePartition = geom::KMeansClustering>
class ParticleMesh final {
public:
  explicit ParticleMesh(Search search = {}, FaceSearch face_search = {}, Partition partition = {}, InterfacePartition interface_partition = {})
      : search_{std::move(search)}, face_search_{std::move(face_search)}, partition_{std::move(partition)}, interface_partition_{std::move(interface_partition)} {}
  auto search_hint() const noexcept -> double { if constexpr (requires { search_.size_hint; }) return double(search_.size_hint); else return 0.0; }
  auto face_search_hint() const noexcept -> double { if constexpr (requires { face_search_.size_hint; }) return double(face_search_.size_hint); else return 0.0; }

  /// Neighbours of particle `a`, ascending, `a` itself included (particle_mesh.hpp:67-72, 137-147).
  template<class Array>
  auto operator[](ParticleView<Array> a) {
    if (!valid_) { a.array().neighbors_(offsets_, cols_); valid_ = true; }
    const auto row = std::span<const std::uint64_t>{cols_}.subspan(offsets_[a.index()], offsets_[a.index() + 1] - offsets_[a.index()]);
    Array* arr = &a.array();
    return row | std::views::transform([arr](std::uint64_t b) { return ParticleView<Array>{*arr, std::size_t(b)}; });
  }
  /// Boundary faces adjacent to particle `a`, ascending: pairs of the face geometry and the
  /// tuple of its vertex (fixed) particles (particle_mesh.hpp:74-82). With C++23 also `mesh[domain, a]`.
  template<class Domain, class Array>
  auto faces(const Domain& domain, ParticleView<Array> a) {
    if (!fvalid_) { a.array().face_neighbors_(foffsets_, fcols_); fvalid_ = true; }
    const auto row = std::span<const std::uint64_t>{fcols_}.subspan(foffsets_[a.index()], foffsets_[a.index() + 1] - foffsets_[a.index()]);
    Array* arr = &a.array();
    const Domain* dom = &domain;
    return row | std::views::transform([arr, dom](std::uint64_t f) {
             const auto& fv = dom->face_verts(std::size_t(f));
             return [&]<std::size_t... I>(std::index_sequence<I...>) {
               return std::pair{dom->face(std::size_t(f)), std::tuple{arr->fixed_at_(fv[I])...}};
             }(std::make_index_sequence<std::tuple_size_v<std::remove_cvref_t<decltype(fv)>>>{});
           });
  }
#if defined(__cpp_multidimensional_subscript)
  template<class Domain, class Array>
  auto operator[](const Domain& domain, ParticleView<Array> a) { return faces(domain, a); }
#endif
  /// Unique pairs (a, b), b < a, of adjacent particles in row order (particle_mesh.hpp:85-92, 228-240).
  template<class Array>
  auto pairs(Array& particles) {
    if (!valid_) { particles.neighbors_(offsets_, cols_); valid_ = true; }
    if (!pvalid_) {
      pairs_.clear();
      for (std::size_t a = 0; a + 1 < offsets_.size(); ++a)
        for (std::uint64_t k = offsets_[a]; k < offsets_[a + 1] && cols_[k] < a; ++k) pairs_.emplace_back(a, std::size_t(cols_[k]));
      pvalid_ = true;
    }
    Array* arr = &particles;
    return std::views::all(pairs_) | std::views::transform([arr](const std::pair<std::size_t, std::size_t>& ab) {
             return std::tuple{ParticleView<Array>{*arr, ab.first}, ParticleView<Array>{*arr, ab.second}};
           });
  }
  /// The same pairs grouped into blocks (particle_mesh.hpp:95-105). The reference colours them
  /// into 2T + 1 blocks so that T host threads never touch one particle at a time; the GPU pair
  /// sums are gathers and need no colouring, so there is one block.
  template<class Array>
  auto block_pairs(Array& particles) { return std::views::single(pairs(particles)); }
  void invalidate_() noexcept { valid_ = fvalid_ = pvalid_ = false; }
private:
  Search search_;
  FaceSearch face_search_;
  Partition partition_;
  InterfacePartition interface_partition_;
  bool valid_ = false, fvalid_ = false, pvalid_ = false;
  std::vector<std::uint64_t> offsets_, cols_, foffsets_, fcols_;
  std::vector<std::pair<std::size_t, std::size_t>> pairs_;
};
template<class S, class F, class P, class I> ParticleMesh(S, F, P, I) -> ParticleMesh<S, F, P, I>;
template<class S, class F> ParticleMesh(S, F) -> ParticleMesh<S, F>;

}  // namespace sph
}  // namespace tit
