// 2-D dam break on the B200 path through the C++ facade: the case of the
// reference's titwcsph executable (/root/reference/source/titwcsph/wcsph.cpp)
// set up and run with the reference's API calls — every call below exists under
// the same name in tit/sph and tit/geom, and INTEGRATION.md §1 shows the two-line
// change that points the reference's own driver at this library.
//
//   case (wcsph.cpp:37-55)      column 2H x H of water, H = 0.6, in a tank 5.366H x 4H;
//                               dr = H / n_col (80 in the reference), h = 2 dr, m = rho_0 dr^2,
//                               c_0 = 20 sqrt(g H), mu = 1e-3, Tait EOS, Wendland C4 kernel
//   walls (:57-81)              four segments tessellated to dr; a second, un-tessellated
//                               copy with the opposite orientation for the containment test
//   particles (:103-142)        fluid on the lattice dr (i + 1, j + 1), one fixed particle per
//                               wall vertex, hydrostatic density from the series solution of
//                               the pressure Poisson problem
//   loop (:157-193)             SSPRK3 steps until t sqrt(g / H) = 10, a frame of all fields
//                               into particles.ttdb at the start and every 100 steps
//
// Extras of this driver: resolution, step limit and output paths come from argv,
// the final r, v, rho can be dumped as raw doubles (tests/test_facade.py), and
// between output frames only the state is published (ParticleArray::publish).
//
//   dam_break_2d [n_col=80] [max_steps=0: run to the end] [dump.bin|-] [particles.ttdb|-]
// A negative n_col stores the set-up of |n_col| (surfaces, initial particles) as one
// frame of the database and stops before the first GPU call.
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <numbers>
#include <optional>
#include <string_view>
#include <vector>

#include "tit_b200/sph.hpp"

namespace {

using namespace tit;
using namespace tit::sph;
using Real = float64_t;
using Point = Vec<Real, 2>;

struct DamBreak {
  int n_col;
  Real H = 0.6, L = 2 * H;                        // water column
  Real pool_width = 5.366 * H, pool_height = 4.0 * H;
  Real g = 9.81, rho_0 = 1000.0, mu = 0.001;
  Real dr = H / Real(n_col), h_0 = 2.0 * dr, m_0 = rho_0 * std::pow(dr, 2), cs_0 = 20 * std::sqrt(g * H);

  /// The tank as a closed polyline through `corners`, in that order.
  static auto polyline(std::initializer_list<Point> corners) -> geom::Surface<Point> {
    geom::Surface<Point> surface;
    for (const auto& corner : corners) surface.append_vert(corner);
    for (std::size_t k = 0; k < corners.size(); ++k) surface.append_face({k, (k + 1) % corners.size()});
    return surface;
  }
  /// Walls, starting at the top-left corner, split into pieces no longer than dr.
  auto walls() const { return geom::tessellate(polyline({{0.0, pool_height}, {pool_width, pool_height}, {pool_width, 0.0}, {0.0, 0.0}}), dr); }
  /// The same rectangle the other way round: winding number one inside.
  auto interior() const { return polyline({{0.0, 0.0}, {pool_width, 0.0}, {pool_width, pool_height}, {0.0, pool_height}}); }

  /// Hydrostatic density at (x, y) of the column: pressure from the first 50 odd
  /// terms of the series solution of the Poisson problem, then the linearised EOS.
  auto hydrostatic_density(Real x, Real y) const -> Real {
    auto p = rho_0 * g * (H - y);
    for (std::size_t k = 1; k < 100; k += 2) {
      const auto k_pi = static_cast<Real>(k) * std::numbers::pi_v<Real>;
      p -= 8 * rho_0 * g * H / pow2(k_pi) * (std::exp(k_pi * (x - L) / (2 * H)) * std::cos(k_pi * y / (2 * H)));
    }
    return rho_0 + p / pow2(cs_0);
  }
};

auto run(int argc, char** argv) -> int {
  const auto arg = [&](int k, const char* fallback) { return argc > k ? argv[k] : fallback; };
  const int n_col_arg = std::atoi(arg(1, "80"));
  const DamBreak cfg{std::abs(n_col_arg)};
  const std::size_t max_steps = std::strtoull(arg(2, "0"), nullptr, 10);
  const std::string_view dump = arg(3, "-"), database = arg(4, "./particles.ttdb");

  const auto domain = cfg.walls();
  const auto inside = cfg.interior();
  const auto containment = geom::MakeFastWinding<Real>{}(inside);
  const FluidEquations equations{cfg.g, cfg.mu, domain, containment, TaitEquationOfState{cfg.cs_0, cfg.rho_0}, SixthOrderWendlandKernel{}};
  const SSPRKIntegrator time_integrator{equations, SSPRKOrder::three};
  ParticleArray particles{Space<Real, 2>{}, time_integrator};

  const auto columns = int(std::round(cfg.L / cfg.dr)), rows = int(std::round(cfg.H / cfg.dr));
  particles.reserve(std::size_t(columns) * rows + domain.num_verts());
  for (auto i = 0; i < columns; ++i)
    for (auto j = 0; j < rows; ++j) r[particles.append(ParticleType::fluid)] = cfg.dr * Vec{i + Real{1.0}, j + Real{1.0}};
  for (std::size_t i = 0; i < domain.num_verts(); ++i) r[particles.append(ParticleType::fixed)] = domain.vert(i);
  h[particles] = cfg.h_0;
  for (const auto a : particles.all()) {
    m[a] = cfg.m_0;
    rho[a] = a.has_type(ParticleType::fixed) ? cfg.rho_0 : cfg.hydrostatic_density(r[a][0], r[a][1]);
  }

  if (n_col_arg < 0) {
    data::Storage setup{database};
    const auto frame = setup.create_series("setup").create_frame(0.0);
    const auto faces_of = [](const geom::Surface<Point>& surface) {
      std::vector<Vec<std::uint64_t, 2>> faces;
      for (const auto& f : surface.face_verts()) faces.emplace_back(f[0], f[1]);
      return faces;
    };
    frame.create_array("verts").write(domain.verts());
    frame.create_array("faces").write(faces_of(domain));
    frame.create_array("containment_verts").write(inside.verts());
    frame.create_array("containment_faces").write(faces_of(inside));
    frame.create_array("r").write(r[particles]);
    frame.create_array("rho").write(rho[particles]);
    frame.create_array("m").write(m[particles]);
    std::printf("setup: %zu fluid + %zu fixed particles, %zu wall faces\n", particles.num_fluid(), particles.num_fixed(), domain.num_faces());
    return 0;
  }

  ParticleMesh mesh{geom::GridSearch{cfg.h_0}, geom::GridFaceSearch{cfg.h_0}, geom::RecursiveInertialBisection{},
                    geom::PixelatedPartition{2 * cfg.h_0, geom::KMeansClustering{}}};
  equations.initialize(mesh, particles);

  // Only the last run is kept in the database.
  std::optional<data::Storage> storage;
  std::optional<data::SeriesView<data::Storage>> series;
  if (database != "-") {
    storage.emplace(database);
    storage->set_max_series(1);
    series = storage->create_series();
    particles.write(0.0, *series);
  }

  const auto started = std::chrono::steady_clock::now();
  const Real time_scale = std::sqrt(cfg.g / cfg.H);
  Real time{};
  std::size_t step = 1;
  for (;; ++step) {
    const Real scaled_time = time * time_scale;
    const bool last = scaled_time >= 10.0 || (max_steps != 0 && step >= max_steps);
    const bool frame = step % 100 == 0 || last;
    particles.publish(frame ? Publish::all : Publish::state);  // derived fields only where a frame follows
    const Real dt = time_integrator.step(mesh, particles);
    if (frame) {
      const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count();
      std::printf("%15zu\t\t%10.5f\t\t%10.5f s/step\t\tdt = %.6e\n", step, double(scaled_time), elapsed / double(step), double(dt));
      // As in the reference the frame carries the time at which its step began; a frame at
      // the very first step would repeat the initial frame's time and is left out.
      if (series && scaled_time > series->last_frame().time()) particles.write(scaled_time, *series);
    }
    if (last) break;
    time += dt;
  }

  // Read-back through the field interface: densities, and the adjacency of particle 0.
  double rho_min = 1e300, rho_max = -1e300;
  for (const auto a : particles.fluid()) rho_min = std::min(rho_min, double(rho[a])), rho_max = std::max(rho_max, double(rho[a]));
  std::size_t degree = 0;
  for ([[maybe_unused]] const auto b : mesh[particles[0]]) ++degree;
  std::printf("steps %zu  particles %zu (%zu fluid)  rho in [%.6f, %.6f]  |mesh[0]| = %zu\n", step, particles.size(), particles.num_fluid(), rho_min, rho_max, degree);

  if (dump != "-") {
    std::FILE* file = std::fopen(dump.data(), "wb");
    if (file == nullptr) throw Exception("cannot open the dump file");
    const auto put = [file](const auto& column) { std::fwrite(column.data(), sizeof(column[0]), column.size(), file); };
    put(r[particles]), put(v[particles]), put(rho[particles]);
    std::fclose(file);
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    tit::par::init();
    return run(argc, argv);
  } catch (const tit::Exception& e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
}
