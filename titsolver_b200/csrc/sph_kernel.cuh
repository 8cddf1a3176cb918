// Device-side smoothing kernel: value, gradient coefficient and the
// semi-analytical wall integrals (flux of W and of its antigradient over a
// boundary segment / triangle clipped by the support sphere).
//
// Replaces, for the GPU path,
//   /root/reference/source/tit/sph/kernel.hpp:142-163  operator(), grad
//   /root/reference/source/tit/sph/kernel.hpp:194-221  flux, antigrad_flux
//   /root/reference/source/tit/sph/kernel.hpp:287-399  unit_segment_integral,
//                                                      unit_triangle_integral
// using the per-kernel recurrences in kernels_gen.cuh (tools/gen_kernels.py).
//
// Boundary faces never move, so their local frames (normal, in-plane axes,
// projected vertices) are precomputed once per surface (FaceFrame) instead of
// being re-derived from the vertices on every evaluation as the reference does.
#pragma once

#include "common.cuh"
#include "kernels_gen.cuh"

namespace titgpu {

template<int D> struct FaceFrame;
// 2-D segment: a, unit normal n = normalize((ba.y, -ba.x)), unit tangent e, length.
template<> struct FaceFrame<2> {
  double a[2], n[2], e[2], len;
  double ctr[2];        // mean of the vertices (r_s of the face)
  double lo[2], hi[2];  // bbox
  unsigned v[2];        // vertex (== fixed particle) indices
};
// 3-D triangle: a, e1 = normalize(ba), n = normalize(cross(ba, ca)),
// e2 = normalize(cross(n_w, e1)); b and c in (e1, e2) coordinates relative to a.
template<> struct FaceFrame<3> {
  double a[3], n[3], e1[3], e2[3];
  double bx, cx, cy;
  double et[3][2], elen[3];  // unit tangent and length of the edges a->b, b->c, c->a in the (e1, e2) frame
  double ctr[3];
  double lo[3], hi[3];
  unsigned v[3];
};

template<int KID>
struct SphKernel {
  using KG = titgpu_gen::KernelGen<KID>;

  template<int D> TIT_HD static constexpr double weight() { return D == 1 ? KG::weight1 : D == 2 ? KG::weight2 : KG::weight3; }

  // W(x) given r = |x| (kernel.hpp:142-151).
  TIT_HD static double value(const Params& P, double rn) { return P.w_val * KG::unit_value(P.hinv * rn); }

  // grad W = c * x with c = w h^-1 w'(q) / |x|; zero for |x| < tiny
  // (normalize(), core/_vec/vec.hpp:663-678).
  TIT_HD static double grad_coef(const Params& P, double d2, double rn) {
    if (d2 < P.tiny2) return 0.0;
    return P.w_val * P.hinv * KG::unit_deriv(P.hinv * rn) / rn;
  }

  // Same with 1 / |x| supplied (the pair loops get it from one rsqrt).
  TIT_HD static double grad_coef_rinv(const Params& P, double rn, double rinv) {
    return P.w_val * P.hinv * KG::unit_deriv(P.hinv * rn) * rinv;
  }

  // The wall integrals are deliberately NOT inlined into their callers: each
  // primitive carries an atan2 and a log1p, and fully unrolled (3 edges x 3
  // pieces x 2 end points x 2 kinds) the wall kernel grew to ~40 k instructions
  // and stalled on instruction fetch. `anti` selects the antigradient
  // primitive at run time so that both kinds share one copy of the code.

  // ---- 2-D: clipped segment integral (kernel.hpp:287-314) ----
  template<int I>
  TIT_HDN static double seg_prim(bool anti, double eta, double z, bool eta_tiny) {
    const double rho = sqrt(fma(z, z, eta * eta));
    const double A = atan2(z, eta);
    const double L = eta_tiny ? 0.0 : copysign(log1p((fabs(z) + z * z / (rho + eta)) / eta), z);
    if (anti) return KG::template seg_antigrad<I>(eta, z, rho, A, L);
    return KG::template seg_flux<I>(eta, z, rho, A, L);
  }
  template<int I>
  TIT_HD static double seg_piece(const Params& P, bool anti, double eta, double z_min, double z_max) {
    const double c = KG::cutoff(I);
    if (eta >= c) return 0.0;
    const double z_clip = sqrt(c * c - eta * eta);
    const double z_lo = fmax(z_min, -z_clip);
    const double z_hi = fmin(z_max, +z_clip);
    if (z_lo >= z_hi) return 0.0;
    const bool et = fabs(eta) <= P.tiny;
    return seg_prim<I>(anti, eta, z_hi, et) - seg_prim<I>(anti, eta, z_lo, et);
  }
  template<int I = 0>
  TIT_HD static double seg_integral(const Params& P, bool anti, double eta, double z_min, double z_max) {
    if constexpr (I >= KG::num_pieces) return 0.0;
    else return seg_piece<I>(P, anti, eta, z_min, z_max) + seg_integral<I + 1>(P, anti, eta, z_min, z_max);
  }

  // ---- 3-D: clipped triangle integral (kernel.hpp:319-399) ----
  template<int I>
  TIT_HDN static double line_prim(double tiny, bool anti, double eta, double delta, double delta_sqr, double beta_sqr, double beta, double z) {
    const double rho = sqrt(fma(z, z, beta_sqr));
    const double A = fabs(delta) <= tiny ? 0.0 : atan2(delta * z * (rho - eta), fma(delta_sqr, rho, z * z * eta));
    const double L = fabs(beta) <= tiny ? 0.0 : copysign(log1p((fabs(z) + z * z / (rho + beta)) / beta), z);
    if (anti) return KG::template tri_antigrad_line<I>(eta, delta, z, rho, A, L);
    return KG::template tri_flux_line<I>(eta, delta, z, rho, A, L);
  }
  // line_prim(z_hi) - line_prim(z_lo) on one line from a single pass of the
  // recurrences (kernels_gen.cuh, *_line_delta) with ONE atan2 and ONE log1p:
  //   A(z) = atan2(y, x) has x > 0, so A in (-pi/2, pi/2) and
  //     A1 - A0 = atan2(y1 x0 - x1 y0, x1 x0 + y1 y0) without wrap-around;
  //   L(z) = asinh(z / beta) = log(w / beta), w = z + rho (z >= 0) or beta^2 / (rho - z),
  //     L1 - L0 = log1p of a single quotient chosen by the signs of z.
  // Same value as the difference of two line_prim() up to rounding (~1e-15).
  template<int I>
  __device__ __noinline__ static double line_prim_delta(double tiny, bool anti, double eta, double delta, double delta_sqr, double beta_sqr, double beta, double z1, double z0) {
    const double rho1 = sqrt(fma(z1, z1, beta_sqr)), rho0 = sqrt(fma(z0, z0, beta_sqr));
    double dA = 0.0, dL = 0.0;
    if (!(fabs(delta) <= tiny)) {
      const double y1 = delta * z1 * (rho1 - eta), x1 = fma(delta_sqr, rho1, z1 * z1 * eta);
      const double y0 = delta * z0 * (rho0 - eta), x0 = fma(delta_sqr, rho0, z0 * z0 * eta);
      dA = atan2(fma(y1, x0, -(x1 * y0)), fma(x1, x0, y1 * y0));
    }
    if (!(fabs(beta) <= tiny)) {
      double num, den;
      if (z0 >= 0.0 && z1 >= 0.0) { den = z0 + rho0; num = (z1 - z0) + (rho1 - rho0); }
      else if (z0 < 0.0 && z1 < 0.0) { den = rho1 - z1; num = (rho0 - rho1) + (z1 - z0); }
      else if (z1 >= 0.0) { const double prod = (z1 + rho1) * (rho0 - z0); den = beta_sqr; num = prod - beta_sqr; }
      else { const double prod = (rho1 - z1) * (z0 + rho0); den = prod; num = beta_sqr - prod; }
      dL = log1p(num / den);
    }
    if (anti) return KG::template tri_antigrad_line_delta<I>(eta, delta, z1, rho1, z0, rho0, dA, dL);
    return KG::template tri_flux_line_delta<I>(eta, delta, z1, rho1, z0, rho0, dA, dL);
  }

  // One edge p0 -> p1 of the projected triangle: the line is cut at its (up to
  // two) intersections with the support circle; pieces inside the circle use the
  // line primitive, pieces outside contribute the sector term.
  template<int I>
  TIT_HDN static double tri_edge(double tiny, bool anti, double eta, double radius_sqr, double sector, double p0x, double p0y, double p1x, double p1y) {
    const double ex = p1x - p0x, ey = p1y - p0y;
    const double len2 = ex * ex + ey * ey;
    if (len2 <= tiny * tiny) return 0.0;
    const double len = sqrt(len2);
    const double tx = ex / len, ty = ey / len;
    const double delta = p0x * ty - p0y * tx;  // det(p0, tangent)
    const double delta_sqr = delta * delta;
    const double beta_sqr = eta * eta + delta_sqr;
    const double beta = sqrt(beta_sqr);
    const double z_start = p0x * tx + p0y * ty;
    const double z_finish = z_start + len;
    // Cut points (kernel.hpp:371-380); an absent cut collapses its piece to zero
    // length, which the `tiny` test below skips just like a short piece.
    double m1 = z_start, m2 = z_start;
    if (radius_sqr > delta_sqr) {
      const double z_clip = sqrt(radius_sqr - delta_sqr);
      if (z_start < -z_clip && -z_clip < z_finish) m1 = -z_clip;
      m2 = m1;
      if (z_start < +z_clip && +z_clip < z_finish) m2 = +z_clip;
    }
    double result = 0.0;
    double z_lo = z_start;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      const double z_hi = k == 0 ? m1 : k == 1 ? m2 : z_finish;
      if (fabs(z_hi - z_lo) > tiny) {
        const double zm = 0.5 * (z_lo + z_hi);
        if (zm * zm + delta_sqr < radius_sqr) {
          result += line_prim<I>(tiny, anti, eta, delta, delta_sqr, beta_sqr, beta, z_hi) - line_prim<I>(tiny, anti, eta, delta, delta_sqr, beta_sqr, beta, z_lo);
        } else {
          result += sector * atan2(delta * (z_hi - z_lo), fma(z_lo, z_hi, delta_sqr));
        }
      }
      if (z_hi != z_lo) z_lo = z_hi;
    }
    return result;
  }
  // The same edge integral arranged for SIMT execution (one edge per lane):
  // the support circle is convex, so at most one of the three pieces of an edge
  // lies inside it. The (cheap) sector terms of the outside pieces are summed
  // first, then every lane makes exactly two calls of the (expensive) line
  // primitive, predicated on the inside piece existing. Piece boundaries, the
  // `tiny` skips and the midpoint classification are those of tri_edge above;
  // only the order of the additions differs.
  template<int I>
  __device__ __forceinline__ static double tri_edge_simt(double tiny, bool anti, double eta, double radius_sqr, double sector, double p0x, double p0y, double tx, double ty, double len) {
    // (tx, ty) and len come precomputed with the face (the reference re-derives
    // them from the projected end points on every evaluation).
    const bool edge_ok = len > tiny;
    const double delta = p0x * ty - p0y * tx;
    const double delta_sqr = delta * delta;
    const double beta_sqr = eta * eta + delta_sqr;
    const double beta = sqrt(beta_sqr);
    const double z_start = p0x * tx + p0y * ty;
    const double z_finish = z_start + len;
    double m1 = z_start, m2 = z_start;
    if (radius_sqr > delta_sqr) {
      const double z_clip = sqrt(radius_sqr - delta_sqr);
      if (z_start < -z_clip && -z_clip < z_finish) m1 = -z_clip;
      m2 = m1;
      if (z_start < +z_clip && +z_clip < z_finish) m2 = +z_clip;
    }
    double result = 0.0, in_lo = 0.0, in_hi = 0.0;
    bool has_in = false;
    double extra = 0.0;  // a second inside piece cannot exist for a convex support; handled for robustness
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double z_lo = k == 0 ? z_start : k == 1 ? m1 : m2;
      const double z_hi = k == 0 ? m1 : k == 1 ? m2 : z_finish;
      if (edge_ok && fabs(z_hi - z_lo) > tiny) {
        const double zm = 0.5 * (z_lo + z_hi);
        if (zm * zm + delta_sqr < radius_sqr) {
          if (!has_in) { has_in = true; in_lo = z_lo; in_hi = z_hi; }
          else extra += line_prim<I>(tiny, anti, eta, delta, delta_sqr, beta_sqr, beta, z_hi) - line_prim<I>(tiny, anti, eta, delta, delta_sqr, beta_sqr, beta, z_lo);
        } else {
          result += sector * atan2(delta * (z_hi - z_lo), fma(z_lo, z_hi, delta_sqr));
        }
      }
    }
    if (has_in) result += line_prim_delta<I>(tiny, anti, eta, delta, delta_sqr, beta_sqr, beta, in_hi, in_lo);
    return result + extra;
  }
  // Unit integral of one EDGE (0: a->b, 1: b->c, 2: c->a) of a face over all
  // kernel pieces; the three edges of a face add up to tri_integral().
  template<int I = 0>
  __device__ __forceinline__ static double tri_edge_integral(const Params& P, bool anti, double eta, double p0x, double p0y, double tx, double ty, double len) {
    if constexpr (I >= KG::num_pieces) return 0.0;
    else {
      const double cut = KG::cutoff(I);
      double r = 0.0;
      if (eta < cut) {
        const double sector = anti ? KG::template tri_antigrad_sector<I>(eta) : KG::template tri_flux_sector<I>(eta);
        r = tri_edge_simt<I>(P.tiny, anti, eta, cut * cut - eta * eta, sector, p0x, p0y, tx, ty, len);
      }
      return r + tri_edge_integral<I + 1>(P, anti, eta, p0x, p0y, tx, ty, len);
    }
  }
  // One edge's share of face_integral<Anti>(P, f, x) (3-D); `anti` selects the
  // primitive at run time so that k_weval holds one copy of the code for both passes.
  __device__ __forceinline__ static double face_edge_integral(const Params& P, const FaceFrame<3>& f, const Vec<3>& x, int edge, bool anti) {
    const double ax = f.a[0] - x[0], ay = f.a[1] - x[1], az = f.a[2] - x[2];
    const double d = -(ax * f.n[0] + ay * f.n[1] + az * f.n[2]) * P.hinv;
    const double pax = (ax * f.e1[0] + ay * f.e1[1] + az * f.e1[2]) * P.hinv;
    const double pay = (ax * f.e2[0] + ay * f.e2[1] + az * f.e2[2]) * P.hinv;
    const double p0x = pax + (edge == 0 ? 0.0 : edge == 1 ? f.bx : f.cx) * P.hinv;
    const double p0y = pay + (edge == 2 ? f.cy : 0.0) * P.hinv;
    const double u = tri_edge_integral(P, anti, fabs(d), p0x, p0y, f.et[edge][0], f.et[edge][1], f.elen[edge] * P.hinv);
    return (anti ? copysign(P.w_anti, d) : P.w_flux) * u;
  }

  template<int I>
  TIT_HD static double tri_piece(const Params& P, bool anti, double eta, const double* a, const double* b, const double* c) {
    const double cut = KG::cutoff(I);
    if (eta >= cut) return 0.0;
    const double radius_sqr = cut * cut - eta * eta;
    const double sector = anti ? KG::template tri_antigrad_sector<I>(eta) : KG::template tri_flux_sector<I>(eta);
    double result = 0.0;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      const double* p0 = k == 0 ? a : k == 1 ? b : c;
      const double* p1 = k == 0 ? b : k == 1 ? c : a;
      result += tri_edge<I>(P.tiny, anti, eta, radius_sqr, sector, p0[0], p0[1], p1[0], p1[1]);
    }
    return result;
  }
  template<int I = 0>
  TIT_HD static double tri_integral(const Params& P, bool anti, double eta, const double* a, const double* b, const double* c) {
    if constexpr (I >= KG::num_pieces) return 0.0;
    else return tri_piece<I>(P, anti, eta, a, b, c) + tri_integral<I + 1>(P, anti, eta, a, b, c);
  }

  // Scalar flux magnitude along the face normal: grad gamma_as = n * flux_n
  // (kernel.hpp:194-206). `Anti` selects the antigradient flux (:209-221), which
  // carries the sign of the wall distance.
  template<bool Anti>
  TIT_HD static double face_integral(const Params& P, const FaceFrame<2>& f, const Vec<2>& x) {
    const double ax = f.a[0] - x[0], ay = f.a[1] - x[1];
    const double d = -(ax * f.n[0] + ay * f.n[1]) * P.hinv;
    const double z_min = (ax * f.e[0] + ay * f.e[1]) * P.hinv;
    const double z_max = z_min + f.len * P.hinv;
    const double u = seg_integral(P, Anti, fabs(d), z_min, z_max);
    if constexpr (Anti) return copysign(P.w_anti, d) * u;
    else return P.w_flux * u;
  }
  template<bool Anti>
  TIT_HD static double face_integral(const Params& P, const FaceFrame<3>& f, const Vec<3>& x) {
    const double ax = f.a[0] - x[0], ay = f.a[1] - x[1], az = f.a[2] - x[2];
    const double d = -(ax * f.n[0] + ay * f.n[1] + az * f.n[2]) * P.hinv;
    double pa[2], pb[2], pc[2];
    pa[0] = (ax * f.e1[0] + ay * f.e1[1] + az * f.e1[2]) * P.hinv;
    pa[1] = (ax * f.e2[0] + ay * f.e2[1] + az * f.e2[2]) * P.hinv;
    pb[0] = pa[0] + f.bx * P.hinv;
    pb[1] = pa[1];
    pc[0] = pa[0] + f.cx * P.hinv;
    pc[1] = pa[1] + f.cy * P.hinv;
    const double u = tri_integral(P, Anti, fabs(d), pa, pb, pc);
    if constexpr (Anti) return copysign(P.w_anti, d) * u;
    else return P.w_flux * u;
  }
};

// Exact sphere/face intersection test of the reference face search
// (geom/segment.hpp:92-112, geom/triangle.hpp:116-186): bbox overlap, then
// closest point on the face within the radius.
TIT_HD bool face_intersects(const FaceFrame<2>& f, const Vec<2>& c, double radius, double radius2, double tiny) {
  for (int d = 0; d < 2; ++d)
    if (!(c[d] - radius <= f.hi[d] && f.lo[d] <= c[d] + radius)) return false;
  // clamp(): a + t * ba with t in [0, 1]; ba = e * len.
  const double px = c[0] - f.a[0], py = c[1] - f.a[1];
  const double len2 = f.len * f.len;
  double qx, qy;
  if (fabs(len2) <= tiny) {
    qx = f.a[0]; qy = f.a[1];
  } else {
    const double bax = f.e[0] * f.len, bay = f.e[1] * f.len;
    const double t = (px * bax + py * bay) / len2;
    if (t < 0.0) { qx = f.a[0]; qy = f.a[1]; }
    else if (t > 1.0) { qx = f.a[0] + bax; qy = f.a[1] + bay; }
    else { qx = f.a[0] + t * bax; qy = f.a[1] + t * bay; }
  }
  const double dx = qx - c[0], dy = qy - c[1];
  return dx * dx + dy * dy <= radius2;
}

TIT_HD bool face_intersects(const FaceFrame<3>& f, const Vec<3>& p, double radius, double radius2, double tiny) {
  for (int d = 0; d < 3; ++d)
    if (!(p[d] - radius <= f.hi[d] && f.lo[d] <= p[d] + radius)) return false;
  // Work in the triangle frame: a = (0,0), b = (bx,0), c = (cx,cy), point
  // (u, v, w) with w the plane distance. Closest point on the triangle follows
  // the region walk of geom/triangle.hpp:137-181 (Ericson), evaluated in-plane.
  const double x = p[0] - f.a[0], y = p[1] - f.a[1], z = p[2] - f.a[2];
  const double u = x * f.e1[0] + y * f.e1[1] + z * f.e1[2];
  const double v = x * f.e2[0] + y * f.e2[1] + z * f.e2[2];
  const double w = x * f.n[0] + y * f.n[1] + z * f.n[2];
  const double bx = f.bx, cx = f.cx, cy = f.cy;
  (void)tiny;
  double qu, qv;
  const double d1 = bx * u, d2 = cx * u + cy * v;
  const double d3 = bx * (u - bx), d4 = cx * (u - bx) + cy * v;
  const double d5 = bx * (u - cx), d6 = cx * (u - cx) + cy * (v - cy);
  const double vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
  if (d1 <= 0.0 && d2 <= 0.0) { qu = 0; qv = 0; }
  else if (d3 >= 0.0 && d4 <= d3) { qu = bx; qv = 0; }
  else if (d6 >= 0.0 && d5 <= d6) { qu = cx; qv = cy; }
  else if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) { const double t = d1 / (d1 - d3); qu = t * bx; qv = 0; }
  else if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { const double t = d2 / (d2 - d6); qu = t * cx; qv = t * cy; }
  else if (va <= 0.0 && d4 >= d3 && d5 >= d6) { const double t = (d4 - d3) / ((d4 - d3) + (d5 - d6)); qu = bx + t * (cx - bx); qv = t * cy; }
  else { const double s = 1.0 / (va + vb + vc); const double tv = vb * s, tw = vc * s; qu = tv * bx + tw * cx; qv = tw * cy; }
  const double du = qu - u, dv = qv - v;
  return du * du + dv * dv + w * w <= radius2;
}

}  // namespace titgpu
