// Stand-ins for the four core-runtime names the reference driver
// (/root/reference/source/titwcsph/wcsph.cpp) uses besides the tit::sph / tit::geom /
// tit::data surface that include/tit_b200/sph.hpp provides: Stopwatch, StopwatchCycle
// (tit/core/time.hpp), log (tit/core/logging.hpp) and TIT_IMPLEMENT_MAIN (tit/core/main.hpp)
// - SURVEY.md section 2 row 9, out of scope of the hot path. `log` is called once per step:
// TIT_SHIM_MAX_STEPS ends the run after that many steps (the driver itself runs ~20 000).
#pragma once
#include <chrono>
#include <cstdlib>
#include <format>
#include <iostream>
namespace tit {
class Stopwatch final {
public:
  auto cycle() const noexcept -> double { return n_ != 0 ? total_ / double(n_) : 0.0; }
  void add(double seconds) noexcept { total_ += seconds; ++n_; }
private:
  double total_ = 0.0;
  std::size_t n_ = 0;
};
class StopwatchCycle final {
public:
  explicit StopwatchCycle(Stopwatch& s) noexcept : s_{&s}, t0_{std::chrono::steady_clock::now()} {}
  ~StopwatchCycle() { s_->add(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count()); }
private:
  Stopwatch* s_;
  std::chrono::steady_clock::time_point t0_;
};
struct ShimStop {};
template<class... Args> void log(std::format_string<Args...> fmt, Args&&... args) {
  static long calls = 0;
  static const long max_steps = std::getenv("TIT_SHIM_MAX_STEPS") != nullptr ? std::atol(std::getenv("TIT_SHIM_MAX_STEPS")) : -1;
  if (max_steps >= 0 && calls++ >= max_steps) throw ShimStop{};
  std::cout << std::format(fmt, std::forward<Args>(args)...) << '\n';
}
}  // namespace tit
#define TIT_IMPLEMENT_MAIN(...) int main(int argc_, char** argv_) { using namespace tit; try { (__VA_ARGS__)(argc_, argv_); } catch (const ShimStop&) {} return 0; }
