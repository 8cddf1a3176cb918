// 3-D dam break on the B200 path through the C++ facade.
//
// The reference ships no 3-D case; this driver keeps the statement order of its
// 2-D one (/root/reference/source/titwcsph/wcsph.cpp:36-193: constants, wall
// surface, containment surface, equations, integrator, particle array, lattice,
// fixed particles on the surface vertices, hydrostatic density, mesh options,
// storage, time loop with a frame every 100 steps) in three dimensions:
//   * column 2H x H x H of water in a closed tank 5.366H x 4H x H;
//   * the walls are a structured mesh (vertex spacing ~ wall_ratio * dr, two
//     right triangles per quad, normals into the tank) rather than a red
//     refinement of 12 large triangles, which would put 4-12 M vertices on the
//     walls of the 10 M-particle case; `geom::tessellate` is
//     available for surfaces that need it;
//   * the containment box is grown by dr/2 so that the wall particles lie
//     strictly inside it (on the surface itself `winding > 0.5` is rounding noise);
//   * optional seeded jitter moves the fluid off the lattice (generic positions
//     for parity runs: faces tangent to a support sphere are evaluated at
//     rounding-noise level by the reference's own triangle integral).
// The same case is built by titsolver_b200.cases.dam_break_3d (used by bench.py
// and the parity tests); tests/test_ttdb.py checks that both give the same mesh
// and, fed the same particles, the same result.
//
//   dam_break_3d [n_col=16] [max_steps=0 (run to t sqrt(g/H) = 10)] [particles.ttdb|-] [wall_ratio=1] [jitter=0]
// A negative n_col writes only the set-up of |n_col| (surfaces and initial
// particles, one frame) and stops before the first GPU call.
#include <array>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <optional>
#include <random>
#include <string_view>
#include <vector>

#include "tit_b200/sph.hpp"

namespace tit::sph::dam_break_3d {
namespace {

using Real = float64_t;
using V3 = Vec<Real, 3>;

/// Triangulated walls of the box [0, ext]: `cells[d]` quads along axis d, vertices
/// shared between walls and numbered in lexicographic (i, j, k) order.
auto box_surface(const std::array<Real, 3>& ext, const std::array<std::int64_t, 3>& cells, bool inward) -> geom::Surface<V3> {
  using Corner = std::array<std::int64_t, 3>;
  std::vector<std::array<Corner, 3>> tris;
  const auto wall = [&](int axis, std::int64_t level, int ua, int va, bool flip) {
    for (std::int64_t iu = 0; iu < cells[ua]; ++iu) {
      for (std::int64_t iv = 0; iv < cells[va]; ++iv) {
        const auto corner = [&](std::int64_t du, std::int64_t dv) {
          Corner c{};
          c[axis] = level, c[ua] = iu + du, c[va] = iv + dv;
          return c;
        };
        // Quad p00 -> p10 (+u) -> p11 -> p01 (+v), its normal is e_u x e_v.
        const Corner p00 = corner(0, 0), p10 = corner(1, 0), p11 = corner(1, 1), p01 = corner(0, 1);
        if (flip) tris.push_back({p00, p11, p10}), tris.push_back({p00, p01, p11});
        else tris.push_back({p00, p10, p11}), tris.push_back({p00, p11, p01});
      }
    }
  };
  wall(2, 0, 0, 1, !inward);        // z = 0, normal +z
  wall(2, cells[2], 0, 1, inward);  // z = ext_z
  wall(1, 0, 2, 0, !inward);        // y = 0, normal +y = e_z x e_x
  wall(1, cells[1], 2, 0, inward);
  wall(0, 0, 1, 2, !inward);        // x = 0, normal +x = e_y x e_z
  wall(0, cells[0], 1, 2, inward);

  std::map<Corner, std::size_t> index;  // lexicographic = the order of the vertex numbers
  for (const auto& t : tris) for (const auto& c : t) index.emplace(c, 0);
  geom::Surface<V3> surface;
  std::size_t next = 0;
  for (auto& [c, id] : index) {
    id = next++;
    surface.append_vert({ext[0] * Real(c[0]) / Real(cells[0]), ext[1] * Real(c[1]) / Real(cells[1]), ext[2] * Real(c[2]) / Real(cells[2])});
  }
  for (const auto& t : tris) surface.append_face({index.at(t[0]), index.at(t[1]), index.at(t[2])});
  return surface;
}

auto sph_main(int argc, char** argv) -> int {
  const int n_col_arg = argc > 1 ? std::atoi(argv[1]) : 16;
  const bool setup_only = n_col_arg < 0;  // -n_col: write the set-up (surfaces, particles) and stop before any GPU call
  const int n_col = std::abs(n_col_arg);
  const std::size_t max_steps = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 0;
  const char* ttdb = argc > 3 ? argv[3] : "./particles.ttdb";
  const Real wall_ratio = argc > 4 ? std::atof(argv[4]) : 1.0;
  const Real jitter = argc > 5 ? std::atof(argv[5]) : 0.0;

  constexpr Real H = 0.6;   // Water column height.
  constexpr Real L = 2 * H; // Water column length.

  constexpr std::array<Real, 3> POOL{5.366 * H, 4.0 * H, 1.0 * H};  // Pool width, height, depth.

  const Real dr = H / Real(n_col);  // Initial particle spacing.
  const auto WATER_M = int(std::round(L / dr));
  const auto WATER_N = int(std::round(H / dr));
  const auto WATER_K = int(std::round(POOL[2] / dr)) - 1;

  constexpr Real g = 9.81;
  constexpr Real rho_0 = 1000.0;
  const Real cs_0 = 20 * std::sqrt(g * H);
  const Real h_0 = 2.0 * dr;
  const Real m_0 = rho_0 * std::pow(dr, 3);
  constexpr Real mu = 0.001;

  // Setup the SPH equations: wall surface ...
  std::array<std::int64_t, 3> wall_cells{};
  for (int d = 0; d < 3; ++d) wall_cells[d] = std::max<std::int64_t>(1, std::int64_t(std::ceil(POOL[d] / (wall_ratio * dr) - 1e-9)));
  const auto domain = box_surface(POOL, wall_cells, /*inward=*/true);

  // ... and another domain for containment tests: 12 triangles, outward normals,
  // half a spacing outside the walls.
  geom::Surface<V3> domain2;
  {
    const auto box = box_surface(POOL, {1, 1, 1}, /*inward=*/false);
    for (const auto& q : box.verts()) {
      V3 p = q;
      for (int d = 0; d < 3; ++d) p[d] = p[d] > 0.0 ? p[d] + 0.5 * dr : p[d] - 0.5 * dr;
      domain2.append_vert(p);
    }
    for (const auto& f : box.face_verts()) domain2.append_face(f);
  }
  const geom::MakeFastWinding<Real> make_winding;
  const auto containment = make_winding(domain2);

  const FluidEquations equations{
      // Constants.
      g,
      mu,
      // Wall boundary.
      domain,
      containment,
      // Weakly compressible equation of state.
      TaitEquationOfState{cs_0, rho_0},
      // C4 Wendland's spline kernel.
      SixthOrderWendlandKernel{},
  };

  // Setup the time integrator.
  const SSPRKIntegrator time_integrator{equations, SSPRKOrder::three};

  // Setup the particles array:
  ParticleArray particles{
      // 3D space.
      Space<Real, 3>{},
      // Set of fields is inferred from the time integrator.
      time_integrator,
  };

  // Generate individual particles.
  std::mt19937_64 random{123};
  const auto shake = [&]() -> Real {  // uniform in [-jitter, jitter) * dr, 53 random bits
    return jitter == 0.0 ? 0.0 : jitter * dr * (2.0 * (Real(random() >> 11) * 0x1.0p-53) - 1.0);
  };
  particles.reserve(std::size_t(WATER_M) * WATER_N * WATER_K + domain.num_verts());
  for (auto i = 0; i < WATER_M; ++i) {
    for (auto j = 0; j < WATER_N; ++j) {
      for (auto k = 0; k < WATER_K; ++k) {
        auto a = particles.append(ParticleType::fluid);
        r[a] = dr * Vec{i + Real{1.0}, j + Real{1.0}, k + Real{1.0}} + Vec{shake(), shake(), shake()};
      }
    }
  }
  for (std::size_t i = 0; i < domain.num_verts(); ++i) {
    auto a = particles.append(ParticleType::fixed);
    r[a] = domain.vert(i);
  }

  // Set global particle constants.
  h[particles] = h_0;
  for (const auto a : particles.all()) {
    m[a] = m_0;
    rho[a] = rho_0;
  }

  // Density hydrostatic initialization (plain column).
  for (const auto a : particles.fluid()) rho[a] = rho_0 + rho_0 * g * (H - r[a][1]) / pow2(cs_0);

  if (setup_only) {
    data::Storage setup{ttdb};
    const auto frame = setup.create_series("setup").create_frame(0.0);
    const auto faces_of = [](const geom::Surface<V3>& surface) {
      std::vector<Vec<std::uint64_t, 3>> faces;
      for (const auto& f : surface.face_verts()) faces.emplace_back(f[0], f[1], f[2]);
      return faces;
    };
    frame.create_array("verts").write(domain.verts());
    frame.create_array("faces").write(faces_of(domain));
    frame.create_array("containment_verts").write(domain2.verts());
    frame.create_array("containment_faces").write(faces_of(domain2));
    frame.create_array("r").write(r[particles]);
    frame.create_array("rho").write(rho[particles]);
    frame.create_array("m").write(m[particles]);
    std::printf("setup: %zu fluid + %zu fixed particles, %zu wall faces\n", particles.num_fluid(), particles.num_fixed(), domain.num_faces());
    return 0;
  }

  // Setup the particle mesh structure.
  ParticleMesh mesh{
      geom::GridSearch{h_0},
      geom::GridFaceSearch{h_0},
      geom::RecursiveInertialBisection{},
      geom::PixelatedPartition{2 * h_0, geom::KMeansClustering{}},
  };

  // Initialize the particles.
  equations.initialize(mesh, particles);

  // Create a data storage to store the particles; only the last run is kept.
  std::optional<data::Storage> storage;
  std::optional<data::SeriesView<data::Storage>> series;
  if (std::string_view{ttdb} != "-") {
    storage.emplace(ttdb);
    storage->set_max_series(1);
    series = storage->create_series();
    particles.write(0.0, *series);
  }

  // Run the simulation.
  Real time{};
  const auto t0 = std::chrono::steady_clock::now();
  std::size_t step = 1;
  for (;; ++step) {
    const auto scaled_time = time * std::sqrt(g / H);
    const auto end_time = 10.0;
    const auto end = scaled_time >= end_time || (max_steps != 0 && step >= max_steps);
    const auto output = (step % 100 == 0) || end;

    // Derived fields are needed only by the step that precedes an output frame.
    particles.publish(output ? Publish::all : Publish::state);
    const Real dt = time_integrator.step(mesh, particles);

    if (output) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::printf("%15zu\t\t%10.5f\t\t%10.5f s/step\t\tdt = %.6e\n", step, double(scaled_time), el / double(step), double(dt));
      if (series && scaled_time > series->last_frame().time()) particles.write(scaled_time, *series);
    }

    if (end) break;
    time += dt;
  }
  std::printf("steps %zu  particles %zu (%zu fluid, %zu wall faces)\n", step, particles.size(), particles.num_fluid(), domain.num_faces());
  return 0;
}

} // namespace
} // namespace tit::sph::dam_break_3d

int main(int argc, char** argv) {
  try {
    tit::par::init();
    return tit::sph::dam_break_3d::sph_main(argc, argv);
  } catch (const tit::Exception& e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
}
