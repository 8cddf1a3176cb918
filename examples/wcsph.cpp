// 2-D dam break on the B200 path through the C++ facade.
//
// This is /root/reference/source/titwcsph/wcsph.cpp restated against
// include/tit_b200/sph.hpp: the set-up (constants, tank surface, lattice,
// hydrostatic density, mesh options) and the time loop keep the reference's
// statements and order (wcsph.cpp:36-193). Differences, all at the edges:
//   * the resolution, the number of steps and an output file come from argv
//     (the reference hard-codes dr = H/80 and runs to t sqrt(g/H) = 10);
//   * the path of the .ttdb storage comes from argv too ("-" = no storage;
//     the reference always writes ./particles.ttdb), and the final r, v, rho
//     can also be dumped as raw doubles;
//   * between output frames only the state is published (particles.publish).
//
//   wcsph [n_col=80] [max_steps=0 (run to the end)] [dump.bin|-] [particles.ttdb|-]
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <numbers>
#include <optional>
#include <string_view>

#include "tit_b200/sph.hpp"

namespace tit::sph::wcsph {
namespace {

template<class Real>
auto sph_main(int argc, char** argv) -> int {
  const int n_col = argc > 1 ? std::atoi(argv[1]) : 80;
  const std::size_t max_steps = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 0;
  const char* dump = argc > 3 && std::string_view{argv[3]} != "-" ? argv[3] : nullptr;
  const char* ttdb = argc > 4 ? argv[4] : "./particles.ttdb";

  constexpr Real H = 0.6;   // Water column height.
  constexpr Real L = 2 * H; // Water column length.

  constexpr Real POOL_WIDTH = 5.366 * H; // Pool width.
  constexpr Real POOL_HEIGHT = 4.0 * H;  // Pool height.

  const Real dr = H / Real(n_col); // Initial particle spacing.
  const auto WATER_M = int(std::round(L / dr));
  const auto WATER_N = int(std::round(H / dr));

  constexpr Real g = 9.81;
  constexpr Real rho_0 = 1000.0;
  const Real cs_0 = 20 * std::sqrt(g * H);
  const Real h_0 = 2.0 * dr;
  const Real m_0 = rho_0 * std::pow(dr, 2);
  constexpr Real mu = 0.001;

  // Setup the SPH equations.
  geom::Surface<Vec<Real, 2>> domain;
  domain.append_vert({0.0, POOL_HEIGHT});
  domain.append_vert({POOL_WIDTH, POOL_HEIGHT});
  domain.append_vert({POOL_WIDTH, 0.0});
  domain.append_vert({0.0, 0.0});
  domain.append_face({0, 1});
  domain.append_face({1, 2});
  domain.append_face({2, 3});
  domain.append_face({3, 0});
  domain = geom::tessellate(domain, dr);

  // Another domain for containment tests.
  geom::Surface<Vec<Real, 2>> domain2;
  domain2.append_vert({0.0, 0.0});
  domain2.append_vert({POOL_WIDTH, 0.0});
  domain2.append_vert({POOL_WIDTH, POOL_HEIGHT});
  domain2.append_vert({0.0, POOL_HEIGHT});
  domain2.append_face({0, 1});
  domain2.append_face({1, 2});
  domain2.append_face({2, 3});
  domain2.append_face({3, 0});
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
      // 2D space.
      Space<Real, 2>{},
      // Set of fields is inferred from the time integrator.
      time_integrator,
  };

  // Generate individual particles.
  particles.reserve(std::size_t(WATER_M) * WATER_N + domain.num_verts());
  for (auto i = 0; i < WATER_M; ++i) {
    for (auto j = 0; j < WATER_N; ++j) {
      auto a = particles.append(ParticleType::fluid);
      r[a] = dr * Vec{i + Real{1.0}, j + Real{1.0}};
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

  // Density hydrostatic initialization.
  for (const auto a : particles.all()) {
    if (a.has_type(ParticleType::fixed)) {
      rho[a] = rho_0;
      continue;
    }

    // Compute pressure from Poisson problem.
    const auto x = r[a][0];
    const auto y = r[a][1];
    auto p_a = rho_0 * g * (H - y);
    for (std::size_t k = 1; k < 100; k += 2) {
      constexpr auto pi = std::numbers::pi_v<Real>;
      const auto k_pi = static_cast<Real>(k) * pi;
      p_a -= 8 * rho_0 * g * H / pow2(k_pi) *
             (std::exp(k_pi * (x - L) / (2 * H)) * std::cos(k_pi * y / (2 * H)));
    }

    // Recalculate density.
    rho[a] = rho_0 + p_a / pow2(cs_0);
  }

  // Setup the particle mesh structure.
  ParticleMesh mesh{
      // Search for the particles using the grid search.
      geom::GridSearch{h_0},
      // Search for the boundary faces using the grid search.
      geom::GridFaceSearch{h_0},
      // Use RIB as the primary partitioning method.
      geom::RecursiveInertialBisection{},
      // Use pixelated K-means as the interface partitioning method.
      geom::PixelatedPartition{2 * h_0, geom::KMeansClustering{}},
  };

  // Initialize the particles.
  equations.initialize(mesh, particles);

  // Create a data storage to store the particles. We'll store only one last
  // run result, all the previous runs will be discarded.
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
      // The frame time of the very first step would repeat the initial frame's.
      if (series && scaled_time > series->last_frame().time()) particles.write(scaled_time, *series);
    }

    if (end) break;
    time += dt;
  }

  // Read the results back through the field interface.
  double rho_min = 1e300, rho_max = -1e300;
  for (const auto a : particles.fluid()) {
    rho_min = std::min(rho_min, double(rho[a]));
    rho_max = std::max(rho_max, double(rho[a]));
  }
  std::size_t nb = 0;
  for ([[maybe_unused]] const auto b : mesh[particles[0]]) ++nb;
  std::printf("steps %zu  particles %zu (%zu fluid)  rho in [%.6f, %.6f]  |mesh[0]| = %zu\n", step, particles.size(), particles.num_fluid(), rho_min, rho_max, nb);
  if (dump != nullptr) {
    std::FILE* f = std::fopen(dump, "wb");
    if (f == nullptr) throw Exception("cannot open the dump file");
    const auto rs = r[particles];
    const auto vs = v[particles];
    const auto ds = rho[particles];
    std::fwrite(rs.data(), sizeof(Vec<Real, 2>), rs.size(), f);
    std::fwrite(vs.data(), sizeof(Vec<Real, 2>), vs.size(), f);
    std::fwrite(ds.data(), sizeof(Real), ds.size(), f);
    std::fclose(f);
  }
  return 0;
}

} // namespace
} // namespace tit::sph::wcsph

int main(int argc, char** argv) {
  try {
    tit::par::init();
    return tit::sph::wcsph::sph_main<tit::float64_t>(argc, argv);
  } catch (const tit::Exception& e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
}
