// tit_b200/core.hpp — the small value types the facade headers share:
// tit::Exception (tit/core/exception.hpp:26-84), tit::Vec / tit::Mat (the subset of
// tit/core/vec.hpp, mat.hpp the driver uses). Packed: Vec<Num, Dim> is Dim numbers.
#pragma once

#include <array>
#include <cmath>
#include <concepts>
#include <cstddef>
#include <stdexcept>

namespace tit {

using float64_t = double;

/// tit/core/exception.hpp:26-84.
class Exception : public std::runtime_error {
public:
  using std::runtime_error::runtime_error;
};

template<class Num> constexpr auto pow2(Num a) noexcept -> Num { return a * a; }

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Vec / Mat: packed (Dim doubles; the reference pads Vec<double,3> to 32 bytes,
// which the C ABI accepts through stride_bytes).

template<class Num, std::size_t Dim>
class Vec final {
public:
  constexpr Vec() noexcept = default;
  template<class... Args>
    requires (sizeof...(Args) == Dim && Dim > 1 && (std::convertible_to<Args, Num> && ...))
  constexpr Vec(Args... qs) noexcept : e_{static_cast<Num>(qs)...} {}
  constexpr explicit Vec(Num q) noexcept { e_.fill(q); }
  constexpr auto operator[](std::size_t i) noexcept -> Num& { return e_[i]; }
  constexpr auto operator[](std::size_t i) const noexcept -> const Num& { return e_[i]; }
  constexpr auto elems() const noexcept -> const std::array<Num, Dim>& { return e_; }
  friend constexpr auto operator+(Vec a, const Vec& b) noexcept -> Vec { for (std::size_t i = 0; i < Dim; ++i) a[i] += b[i]; return a; }
  friend constexpr auto operator-(Vec a, const Vec& b) noexcept -> Vec { for (std::size_t i = 0; i < Dim; ++i) a[i] -= b[i]; return a; }
  friend constexpr auto operator*(Num s, Vec a) noexcept -> Vec { for (auto& q : a.e_) q *= s; return a; }
  friend constexpr auto operator*(Vec a, Num s) noexcept -> Vec { for (auto& q : a.e_) q *= s; return a; }
  friend constexpr auto operator/(Vec a, Num s) noexcept -> Vec { for (auto& q : a.e_) q /= s; return a; }
  friend constexpr auto operator==(const Vec&, const Vec&) noexcept -> bool = default;
private:
  std::array<Num, Dim> e_{};
};
template<class Num, class... Rest> Vec(Num, Rest...) -> Vec<Num, 1 + sizeof...(Rest)>;

template<class Num, std::size_t Dim>
constexpr auto dot(const Vec<Num, Dim>& a, const Vec<Num, Dim>& b) noexcept -> Num {
  Num r = a[0] * b[0];
  for (std::size_t i = 1; i < Dim; ++i) r += a[i] * b[i];
  return r;
}
template<class Num, std::size_t Dim> auto norm(const Vec<Num, Dim>& a) noexcept -> Num { return std::sqrt(dot(a, a)); }

template<class Num, std::size_t Dim>
class Mat final {
public:
  constexpr auto operator[](std::size_t i) noexcept -> Vec<Num, Dim>& { return r_[i]; }
  constexpr auto operator[](std::size_t i) const noexcept -> const Vec<Num, Dim>& { return r_[i]; }
private:
  std::array<Vec<Num, Dim>, Dim> r_{};
};

template<class V> struct vec_traits;
template<class Num, std::size_t Dim> struct vec_traits<Vec<Num, Dim>> { using num = Num; static constexpr std::size_t dim = Dim; };
template<class V> using vec_num_t = typename vec_traits<V>::num;
template<class V> inline constexpr std::size_t vec_dim_v = vec_traits<V>::dim;

}  // namespace tit
