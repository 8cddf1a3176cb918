// Known-answer tests of include/tit_b200/data.hpp: the cases of the reference's
// tit/data/type.test.cpp and tit/data/storage.test.cpp restated as one
// self-checking program (no doctest here), plus the ParticleArray::write path
// of the facade without a GPU (host columns only).
//
//   test_data <scratch dir> [file written by titsolver_b200.ttdb to read back]
//
// Prints "ok <n checks>" and exits 0, or the first failed check and exits 1.
// The file <scratch dir>/particles_cpp.ttdb is left behind for the Python
// reader (tests/test_ttdb.py).
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <numbers>
#include <set>
#include <string>
#include <vector>

#include "tit_b200/sph.hpp"

namespace {

using namespace tit;
namespace fs = std::filesystem;

int n_checks = 0;

#define CHECK(...)                                                               \
  do {                                                                           \
    ++n_checks;                                                                  \
    if (!(__VA_ARGS__)) {                                                              \
      std::fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #__VA_ARGS__); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (false)

/// The expression throws tit::Exception whose message contains `text`.
#define CHECK_THROWS_MSG(expr, text)                                             \
  do {                                                                           \
    ++n_checks;                                                                  \
    bool thrown_ = false;                                                        \
    try { (void)(expr); } catch (const Exception& e) {                           \
      thrown_ = std::string{e.what()}.find(text) != std::string::npos;           \
      if (!thrown_) std::fprintf(stderr, "%s:%d: message was: %s\n", __FILE__, __LINE__, e.what()); \
    }                                                                            \
    if (!thrown_) {                                                              \
      std::fprintf(stderr, "%s:%d: %s did not throw \"%s\"\n", __FILE__, __LINE__, #expr, text); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (false)

template<class Views, class View>
auto same(const Views& got, std::initializer_list<View> want) -> bool {
  return got.size() == want.size() && std::equal(got.begin(), got.end(), want.begin());
}

// ~~ type.test.cpp ~~
void test_types() {
  const data::Kind kind{data::Kind::ID::float32};
  CHECK(kind.id() == data::Kind::ID::float32);
  CHECK(std::string{kind.name()} == "float32_t");
  CHECK(kind.width() == 4);
  CHECK_THROWS_MSG(data::Kind{data::Kind::ID{137}}, "Invalid data kind ID: 137.");
  const std::size_t widths[] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 8};
  for (unsigned k = 0; k < 10; ++k) CHECK(data::Kind{data::Kind::ID(k)}.width() == widths[k]);

  CHECK(data::kind_of<std::int16_t>.id() == data::Kind::ID::int16);
  CHECK(data::kind_of<float32_t>.id() == data::Kind::ID::float32);
  CHECK(data::kind_of<std::uint64_t>.id() == data::Kind::ID::uint64);

  {
    const data::Type type{data::kind_of<float32_t>};
    CHECK(type.kind() == data::kind_of<float32_t> && type.rank() == data::Rank::scalar && type.dim() == 1 && type.width() == 4);
    CHECK(type.name() == "float32_t");
  }
  {
    const data::Type type{data::kind_of<float64_t>, data::Rank::vector, 2};
    CHECK(type.kind() == data::kind_of<float64_t> && type.rank() == data::Rank::vector && type.dim() == 2 && type.width() == 2 * 8);
    CHECK(type.name() == "Vec<float64_t, 2>");
    CHECK(type.id() == 131338);  // the id found in the reference's fixture
  }
  {
    const data::Type type{data::kind_of<std::int16_t>, data::Rank::matrix, 3};
    CHECK(type.kind() == data::kind_of<std::int16_t> && type.rank() == data::Rank::matrix && type.dim() == 3 && type.width() == 3 * 3 * 2);
    CHECK(type.name() == "Mat<int16_t, 3>");
  }
  CHECK_THROWS_MSG(data::Type(data::kind_of<float32_t>, data::Rank{137}, 3), "Invalid data type rank: 137.");
  CHECK_THROWS_MSG(data::Type(data::kind_of<float32_t>, data::Rank::vector, 0), "Dimensionality must be positive, but is 0.");
  CHECK_THROWS_MSG(data::Type(data::kind_of<float32_t>, data::Rank::scalar, 2), "Dimensionality of a scalar must be 1, but is 2.");
  CHECK((data::type_of<Mat<float32_t, 3>>.id() == 0x030209));
  CHECK((data::Type{0x030209} == data::type_of<Mat<float32_t, 3>>));
  CHECK_THROWS_MSG(data::Type{0x1337}, "Invalid");
  CHECK((data::type_of<Vec<std::int16_t, 7>>.rank() == data::Rank::vector && data::type_of<Vec<std::int16_t, 7>>.dim() == 7));
  CHECK((data::type_of<Mat<float64_t, 5>>.rank() == data::Rank::matrix && data::type_of<Mat<float64_t, 5>>.dim() == 5));
  CHECK(data::type_of<float32_t>.rank() == data::Rank::scalar);
}

// ~~ storage.test.cpp: data::Storage ~~
void test_open(const fs::path& dir) {
  const fs::path file = dir / "test.ttdb";
  {
    const data::Storage storage{":memory:"};
    CHECK(storage.path().empty());
  }
  fs::remove(file);
  {
    const data::Storage storage{file};
    CHECK(fs::exists(file));
    CHECK(storage.path().filename() == file.filename());
  }
  {
    const data::Storage storage{file};  // open existing
    CHECK(storage.path().filename() == file.filename());
  }
  {
    data::Storage storage{file, /*read_only=*/true};
    CHECK_THROWS_MSG(storage.create_series_id("test"), "attempt to write a readonly database");
  }
  CHECK_THROWS_MSG(data::Storage{"/invalid/path/to/file.ttdb"}, "unable to open database file");
  const fs::path junk = dir / "junk.ttdb";
  std::ofstream{junk} << "definitely not an SQLite database, just some text long enough to look like a header";
  CHECK_THROWS_MSG(data::Storage{junk}, "file is not a database");
}

// ~~ storage.test.cpp: data::SeriesView ~~
void test_series() {
  {
    const data::Storage storage{":memory:"};
    CHECK(storage.num_series() == 0);
    CHECK(storage.series().empty());  // through the const overloads
  }
  {
    data::Storage storage{":memory:"};
    CHECK(storage.max_series() >= 3);
    const auto s1 = storage.create_series("1");
    CHECK(storage.check_series(s1) && s1 == data::SeriesID{1} && s1.name() == "1" && storage.num_series() == 1);
    CHECK(same(storage.series(), {s1}) && storage.last_series() == s1);
    const auto s2 = storage.create_series("2");
    CHECK(storage.check_series(s2) && s2 == data::SeriesID{2} && s2.name() == "2" && storage.num_series() == 2);
    CHECK(same(storage.series(), {s1, s2}) && storage.last_series() == s2);
    const auto s3 = storage.create_series("3");
    CHECK(s3 == data::SeriesID{3} && s3.name() == "3" && storage.num_series() == 3);
    CHECK(same(storage.series(), {s1, s2, s3}) && storage.last_series() == s3);
    CHECK(storage.series(0) == s1 && storage.series(1) == s2 && storage.series(2) == s3);
    CHECK_THROWS_MSG(storage.series(3), "out of bounds");
    const data::Storage& view = storage;
    const data::SeriesView<const data::Storage> c2 = s2;  // mutable -> const handle
    CHECK(view.series(1) == c2 && view.last_series().name() == "3");
  }
  {  // more series than the maximum: the oldest go
    data::Storage storage{":memory:"};
    storage.set_max_series(3);
    CHECK(storage.max_series() == 3);
    const auto s1 = storage.create_series("1"), s2 = storage.create_series("2"), s3 = storage.create_series("3");
    CHECK(same(storage.series(), {s1, s2, s3}));
    const auto s4 = storage.create_series("4");
    CHECK(storage.check_series(s4) && !storage.check_series(s1) && same(storage.series(), {s2, s3, s4}));
    const auto s5 = storage.create_series("5");
    CHECK(storage.check_series(s5) && !storage.check_series(s2) && same(storage.series(), {s3, s4, s5}));
  }
  {  // decrease / increase the maximum
    data::Storage storage{":memory:"};
    storage.set_max_series(3);
    const auto s1 = storage.create_series("1"), s2 = storage.create_series("2"), s3 = storage.create_series("3");
    storage.set_max_series(2);
    CHECK(storage.max_series() == 2 && same(storage.series(), {s2, s3}) && !storage.check_series(s1));
    storage.set_max_series(5);
    CHECK(storage.max_series() == 5 && same(storage.series(), {s2, s3}));
    const auto s4 = storage.create_series("4"), s5 = storage.create_series("5"), s6 = storage.create_series("6");
    CHECK(same(storage.series(), {s2, s3, s4, s5, s6}));
    CHECK_THROWS_MSG(storage.set_max_series(0), "must be positive");
  }
  {  // delete: ids are not reused
    data::Storage storage{":memory:"};
    storage.set_max_series(3);
    const auto s1 = storage.create_series("1"), s2 = storage.create_series("2"), s3 = storage.create_series("3");
    storage.delete_series(s2);
    CHECK(!storage.check_series(s2) && same(storage.series(), {s1, s3}));
    const auto s4 = storage.create_series("4");
    CHECK(storage.check_series(s4) && !(s4 == s2) && same(storage.series(), {s1, s3, s4}));
  }
}

// ~~ storage.test.cpp: data::FrameView ~~
void test_frames() {
  data::Storage storage{":memory:"};
  const auto series = storage.create_series("");
  CHECK(series.num_frames() == 0);
  const auto f1 = series.create_frame(0.0);
  CHECK(storage.check_frame(f1) && f1 == data::FrameID{1} && f1.time() == 0.0 && series.num_frames() == 1);
  CHECK(same(series.frames(), {f1}) && series.last_frame() == f1);
  const auto f2 = series.create_frame(1.0);
  CHECK(f2 == data::FrameID{2} && f2.time() == 1.0 && same(series.frames(), {f1, f2}) && series.last_frame() == f2);
  const auto f3 = series.create_frame(2.0);
  CHECK(f3 == data::FrameID{3} && f3.time() == 2.0 && series.num_frames() == 3 && same(series.frames(), {f1, f2, f3}));
  CHECK(series.frame(0) == f1 && series.frame(1) == f2 && series.frame(2) == f3);
  CHECK_THROWS_MSG(series.frame(3), "out of bounds");
  CHECK_THROWS_MSG(series.create_frame(2.0), "greater than the last frame time");

  // frames are not shared between series
  const auto other = storage.create_series("");
  const auto g1 = other.create_frame(0.0), g2 = other.create_frame(1.0), g3 = other.create_frame(2.0);
  CHECK(same(other.frames(), {g1, g2, g3}) && same(series.frames(), {f1, f2, f3}));
  const std::set<data::FrameID> all{f1, f2, f3, g1, g2, g3};
  CHECK(all.size() == 6);

  storage.delete_frame(f2);
  CHECK(!storage.check_frame(f2) && same(series.frames(), {f1, f3}));
  const auto f4 = series.create_frame(3.0);
  CHECK(storage.check_frame(f4) && f4 == data::FrameID{7} && same(series.frames(), {f1, f3, f4}));

  storage.delete_series(other);  // cascades
  CHECK(!storage.check_series(other) && !storage.check_frame(g1) && !storage.check_frame(g2) && !storage.check_frame(g3) && storage.check_frame(f1));
}

// ~~ storage.test.cpp: data::ArrayView ~~
void test_arrays() {
  data::Storage storage{":memory:"};
  const auto series = storage.create_series("");
  const auto frame = series.create_frame(0.0);
  CHECK(frame.num_arrays() == 0);
  const auto a1 = frame.create_array("array_1");
  a1.write(std::vector{std::numbers::pi});
  CHECK(storage.check_array(a1) && a1 == data::ArrayID{1} && a1.name() == "array_1" && a1.type() == data::type_of<float64_t> && a1.size() == 1);
  CHECK(a1.read<float64_t>() == std::vector{std::numbers::pi});
  CHECK(frame.num_arrays() == 1 && same(frame.arrays(), {a1}));
  const auto a2 = frame.create_array("array_2");
  const float32_t e = std::numbers::e_v<float32_t>;
  a2.write(data::type_of<float32_t>, std::as_bytes(std::span<const float32_t>{&e, 1}));
  CHECK(a2 == data::ArrayID{2} && a2.name() == "array_2" && a2.type() == data::type_of<float32_t> && a2.size() == 1);
  CHECK(a2.read<float32_t>() == std::vector{e});
  CHECK(frame.num_arrays() == 2 && same(frame.arrays(), {a1, a2}));
  CHECK_THROWS_MSG(a2.read<float64_t>(), "Type mismatch");
  CHECK_THROWS_MSG(frame.create_array("array_2"), "already exists");
  CHECK_THROWS_MSG(frame.create_array(""), "must not be empty");

  // find
  CHECK(frame.find_array("array_1") == a1 && frame.find_array("array_2") == a2 && !frame.find_array("does_not_exist"));

  // update
  a1.write(std::vector{std::numbers::phi, std::numbers::sqrt3});
  CHECK(a1.size() == 2 && (a1.read<float64_t>() == std::vector{std::numbers::phi, std::numbers::sqrt3}));

  // vectors, matrices, an empty array, raw bytes
  const std::vector<Vec<float64_t, 3>> vs{{1.0, 2.0, 3.0}, {4.0, 5.0, 6.0}};
  const auto av = frame.create_array("vectors");
  av.write(vs);
  CHECK(av.type().id() == ((9 + 1) | 1 << 8 | 3 << 16) && av.size() == 2 && av.read().size() == 48);
  CHECK((av.read<Vec<float64_t, 3>>() == vs));
  std::vector<Mat<float64_t, 2>> ms(3);
  for (std::size_t k = 0; k < ms.size(); ++k) { ms[k][0] = {double(k), 1.0}; ms[k][1] = {2.0, -double(k)}; }
  const auto am = frame.create_array("matrices");
  am.write(ms);
  const auto back = am.read<Mat<float64_t, 2>>();
  CHECK(am.type().name() == "Mat<float64_t, 2>" && back.size() == 3 && back[2][1][1] == -2.0 && back[1][0][0] == 1.0);
  const auto ae = frame.create_array("empty");
  ae.write(std::vector<std::uint64_t>{});
  CHECK(ae.size() == 0 && ae.type() == data::type_of<std::uint64_t> && ae.read<std::uint64_t>().empty());
  std::vector<std::byte> three(3);
  CHECK_THROWS_MSG(ae.write(data::type_of<std::uint16_t>, three), "Data size mismatch");
  // something compressible and something large enough to span several zstd blocks
  std::vector<float64_t> big(1 << 18);
  for (std::size_t k = 0; k < big.size(); ++k) big[k] = std::sin(0.001 * double(k)) * double(k % 97);
  const auto ab = frame.create_array("big");
  ab.write(big);
  CHECK(ab.read<float64_t>() == big);

  // delete: ids are not reused; deleting the frame cascades
  storage.delete_array(a1);
  CHECK(!storage.check_array(a1) && frame.arrays().front() == a2);
  const auto a3 = frame.create_array("array_3");
  a3.write(std::vector{std::numbers::phi});
  CHECK(storage.check_array(a3) && a3 == data::ArrayID{7});
  storage.delete_frame(frame);
  CHECK(!storage.check_frame(frame) && !storage.check_array(a2) && !storage.check_array(a3));
}

// ~~ the facade's ParticleArray::write (particle_array.hpp:165-172), host columns only ~~
void test_particle_array_write(const fs::path& file) {
  using namespace tit::sph;
  fs::remove(file);
  const geom::Surface<Vec<double, 2>> none;
  const auto winding = geom::make_exact_winding(none);
  const FluidEquations equations{9.81, 1e-3, none, winding, TaitEquationOfState{10.0, 1000.0}, QuarticWendlandKernel{}};
  const SSPRKIntegrator integrator{equations, SSPRKOrder::three};
  ParticleArray particles{Space<double, 2>{}, integrator};
  for (int i = 0; i < 7; ++i) {
    const auto a = particles.append(i < 5 ? ParticleType::fluid : ParticleType::fixed);
    r[a] = Vec{0.25 * i, 1.0 - 0.125 * i};
    v[a] = Vec{double(i), -double(i)};
    rho[a] = 1000.0 + i;
    m[a] = 0.5;
    L[a][0] = Vec{1.0, double(i)};
    L[a][1] = Vec{-double(i), 2.0};
  }
  data::Storage storage{file};
  storage.set_max_series(1);
  const auto series = storage.create_series();
  particles.write(0.0, series);
  rho[particles[0]] = 999.0;
  particles.write(0.5, series);
  CHECK(series.num_frames() == 2 && series.last_frame().time() == 0.5);
  const auto frame = series.frame(0);
  CHECK(frame.num_arrays() == std::size_t(num_varying_fields));
  std::size_t k = 0;
  for (const auto& array : frame.arrays()) {  // the field-set order of fluid_equations.hpp:41-48
    CHECK(array.name() == varying_field_names[k]);
    CHECK(array.size() == 7);
    ++k;
  }
  CHECK(frame.find_array("r")->type() == (data::type_of<Vec<double, 2>>));
  CHECK(frame.find_array("L")->type() == (data::type_of<Mat<double, 2>>));
  CHECK(frame.find_array("rho")->type() == data::type_of<double>);
  const auto rs = frame.find_array("r")->read<Vec<double, 2>>();
  CHECK((rs[3] == Vec{0.75, 0.625}));
  CHECK(frame.find_array("rho")->read<double>()[0] == 1000.0);
  CHECK(series.frame(1).find_array("rho")->read<double>()[0] == 999.0);
  CHECK(frame.find_array("L")->read<Mat<double, 2>>()[6][1][0] == -6.0);
}

// ~~ a database written by titsolver_b200/ttdb.py ~~
void test_read_python(const fs::path& file) {
  const data::Storage storage{file, /*read_only=*/true};
  CHECK(storage.num_series() == 1);
  const auto series = storage.last_series();
  CHECK(series.name() == "from python" && series.num_frames() == 2);
  const auto frame = series.last_frame();
  CHECK(frame.time() == 1.5);
  const auto rs = frame.find_array("r")->read<Vec<double, 3>>();
  CHECK(rs.size() == 5 && (rs[4] == Vec{12.0, 13.0, 14.0}));
  const auto ls = frame.find_array("L")->read<Mat<double, 3>>();
  CHECK(ls.size() == 5 && ls[1][2][0] == 15.0);
  const auto ids = frame.find_array("parinfo")->read<std::uint64_t>();
  CHECK(ids.size() == 5 && ids[3] == 3);
  CHECK(frame.find_array("rho")->type() == data::type_of<float32_t>);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: test_data <scratch dir> [python.ttdb]\n"); return 2; }
  try {
    const fs::path dir{argv[1]};
    test_types();
    test_open(dir);
    test_series();
    test_frames();
    test_arrays();
    test_particle_array_write(dir / "particles_cpp.ttdb");
    if (argc > 2) test_read_python(argv[2]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "unexpected exception: %s\n", e.what());
    return 1;
  }
  std::printf("ok %d checks\n", n_checks);
  return 0;
}
