// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// Restatement of the reference smoothing-kernel wrapper
// (/root/reference/source/tit/sph/kernel.hpp:125-221 for value/grad/
// width_deriv/antigrad/flux/antigrad_flux, :287-410 for the clipped segment and
// triangle integrals). The per-kernel polynomials come from
// kernels_gen_oracle.h (tools/gen_kernels.py, oracle form).
#pragma once

#include "kernels_gen_oracle.h"
#include "oracle_math.h"

namespace orc {

template<class KG>
struct Kernel {
  template<int D> static constexpr double weight() {
    if constexpr (D == 1) return KG::weight1;
    else if constexpr (D == 2) return KG::weight2;
    else return KG::weight3;
  }
  template<int D> static double moment(double q) {
    if constexpr (D == 1) return KG::unit_moment1(q);
    else if constexpr (D == 2) return KG::unit_moment2(q);
    else return KG::unit_moment3(q);
  }
  static double radius(double h) { return KG::unit_radius * h; }

  template<int D> static double hpow(double hi) {
    double r = hi;
    for (int i = 1; i < D; ++i) r *= hi;
    return r;
  }

  // kernel.hpp:142-151.
  template<int D> static double value(const Vec<D>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<D>() * hpow<D>(hi);
    const double q = hi * norm(x);
    return w * KG::unit_value(q);
  }
  // kernel.hpp:154-163.
  template<int D> static Vec<D> grad(const Vec<D>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<D>() * hpow<D>(hi);
    const double q = hi * norm(x);
    const Vec<D> grad_q = normalize(x) * hi;
    return (w * KG::unit_deriv(q)) * grad_q;
  }
  // kernel.hpp:166-177.
  template<int D> static double width_deriv(const Vec<D>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<D>() * hpow<D>(hi);
    const double dw_dh = -double(D) * w * hi;
    const double q = hi * norm(x);
    const double dq_dh = -q * hi;
    return dw_dh * KG::unit_value(q) + w * KG::unit_deriv(q) * dq_dh;
  }
  // kernel.hpp:181-191.
  template<int D> static Vec<D> antigrad(const Vec<D>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<D>() * hpow<D>(hi);
    const double q = norm(x) * hi;
    return (-1.0 * x) * (w * moment<D>(q) / hpow<D>(q));
  }

  // kernel.hpp:287-314.
  template<class Prim>
  static double unit_segment_integral(double cutoff, double eta, double z_min, double z_max, Prim&& prim) {
    if (eta >= cutoff) return 0.0;
    const double z_clip = std::sqrt(pow2(cutoff) - pow2(eta));
    const double z_lo = std::max(z_min, -z_clip);
    const double z_hi = std::min(z_max, +z_clip);
    if (z_lo >= z_hi) return 0.0;
    const auto eval = [&](double z) {
      const double rho = std::sqrt(pow2(z) + pow2(eta));
      const double A = std::atan2(z, eta);
      const double L = is_tiny(eta) ? 0.0 : std::copysign(std::log1p((std::abs(z) + pow2(z) / (rho + eta)) / eta), z);
      return prim(eta, z, rho, A, L);
    };
    return eval(z_hi) - eval(z_lo);
  }

  // kernel.hpp:319-399.
  template<class Line, class Sector>
  static double unit_triangle_integral(double cutoff, double eta, const Vec<2>& a, const Vec<2>& b, const Vec<2>& c, Line&& line, Sector&& sector) {
    if (eta >= cutoff) return 0.0;
    const double radius_sqr = pow2(cutoff) - pow2(eta);
    const double sector_integral = sector(eta);
    const auto edge_integral = [&](const Vec<2>& p0, const Vec<2>& p1) -> double {
      const Vec<2> edge = p1 - p0;
      const double edge_len_sqr = norm2(edge);
      if (edge_len_sqr <= pow2(tiny)) return 0.0;
      const double edge_len = std::sqrt(edge_len_sqr);
      const Vec<2> tangent = edge / edge_len;
      const double delta = det(p0, tangent);
      const double delta_sqr = pow2(delta);
      const double beta_sqr = pow2(eta) + delta_sqr;
      const double beta = std::sqrt(beta_sqr);
      const auto eval_line = [&](double z) {
        const double rho = std::sqrt(pow2(z) + beta_sqr);
        const double A = is_tiny(delta) ? 0.0 : std::atan2(delta * z * (rho - eta), delta_sqr * rho + pow2(z) * eta);
        const double L = is_tiny(beta) ? 0.0 : std::copysign(std::log1p((std::abs(z) + pow2(z) / (rho + beta)) / beta), z);
        return line(eta, delta, z, rho, A, L);
      };
      const double z_start = dot(p0, tangent);
      const double z_finish = z_start + edge_len;
      double zs[4];
      int nz = 0;
      zs[nz++] = z_start;
      if (radius_sqr > delta_sqr) {
        const double z_clip = std::sqrt(radius_sqr - delta_sqr);
        if (z_start < -z_clip && -z_clip < z_finish) zs[nz++] = -z_clip;
        if (z_start < +z_clip && +z_clip < z_finish) zs[nz++] = +z_clip;
      }
      zs[nz++] = z_finish;
      double result = 0.0;
      for (int i = 0; i + 1 < nz; ++i) {
        const double z_lo = zs[i], z_hi = zs[i + 1];
        if (is_tiny(z_hi - z_lo)) continue;
        if (pow2((z_lo + z_hi) / 2.0) + delta_sqr < radius_sqr) {
          result += eval_line(z_hi) - eval_line(z_lo);
        } else {
          const double arc_angle = std::atan2(delta * (z_hi - z_lo), z_lo * z_hi + delta_sqr);
          result += sector_integral * arc_angle;
        }
      }
      return result;
    };
    return edge_integral(a, b) + edge_integral(b, c) + edge_integral(c, a);
  }

  static double unit_flux(double eta, double z_min, double z_max) {
    double r = 0.0;
    for (int i = 0; i < KG::num_pieces; ++i)
      r += unit_segment_integral(KG::cutoff(i), eta, z_min, z_max, [i](double e, double z, double rho, double A, double L) { return KG::seg_flux(i, e, z, rho, A, L); });
    return r;
  }
  static double unit_antigrad_flux(double eta, double z_min, double z_max) {
    double r = 0.0;
    for (int i = 0; i < KG::num_pieces; ++i)
      r += unit_segment_integral(KG::cutoff(i), eta, z_min, z_max, [i](double e, double z, double rho, double A, double L) { return KG::seg_antigrad(i, e, z, rho, A, L); });
    return r;
  }
  static double unit_flux(double eta, const Vec<2>& a, const Vec<2>& b, const Vec<2>& c) {
    double r = 0.0;
    for (int i = 0; i < KG::num_pieces; ++i)
      r += unit_triangle_integral(
          KG::cutoff(i), eta, a, b, c, [i](double e, double d, double z, double rho, double A, double L) { return KG::tri_flux_line(i, e, d, z, rho, A, L); },
          [i](double e) { return KG::tri_flux_sector(i, e); });
    return r;
  }
  static double unit_antigrad_flux(double eta, const Vec<2>& a, const Vec<2>& b, const Vec<2>& c) {
    double r = 0.0;
    for (int i = 0; i < KG::num_pieces; ++i)
      r += unit_triangle_integral(
          KG::cutoff(i), eta, a, b, c, [i](double e, double d, double z, double rho, double A, double L) { return KG::tri_antigrad_line(i, e, d, z, rho, A, L); },
          [i](double e) { return KG::tri_antigrad_sector(i, e); });
    return r;
  }

  // kernel.hpp:194-206.
  static Vec<2> flux(const Segment& f, const Vec<2>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<2>() * hi;
    const Vec<2> n = f.normal();
    const double d = dot(x - f.a, n) * hi;
    const auto ps = f.project(x);
    return n * (w * unit_flux(std::abs(d), ps[0] * hi, ps[1] * hi));
  }
  static Vec<3> flux(const Triangle& f, const Vec<3>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<3>() * hi;
    const Vec<3> n = f.normal();
    const double d = dot(x - f.a, n) * hi;
    const auto ps = f.project(x);
    return n * (w * unit_flux(std::abs(d), ps[0] * hi, ps[1] * hi, ps[2] * hi));
  }
  // kernel.hpp:209-221.
  static double antigrad_flux(const Segment& f, const Vec<2>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<2>();
    const double d = dot(x - f.a, f.normal()) * hi;
    const auto ps = f.project(x);
    return std::copysign(w, d) * unit_antigrad_flux(std::abs(d), ps[0] * hi, ps[1] * hi);
  }
  static double antigrad_flux(const Triangle& f, const Vec<3>& x, double h) {
    const double hi = 1.0 / h;
    const double w = weight<3>();
    const double d = dot(x - f.a, f.normal()) * hi;
    const auto ps = f.project(x);
    return std::copysign(w, d) * unit_antigrad_flux(std::abs(d), ps[0] * hi, ps[1] * hi, ps[2] * hi);
  }
};

}  // namespace orc
