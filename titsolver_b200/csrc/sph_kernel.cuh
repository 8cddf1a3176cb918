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
  double b[2];          // second vertex (world coordinates: the exact intersection test)
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
// (geom/segment.hpp:92-112, geom/triangle.hpp:116-186, geom/bsphere.hpp:47-53): bbox
// overlap, then the closest point on the face (clamp) within the radius, inclusive. The
// arithmetic is the reference's own, operation by operation and without FMA contraction,
// on the world coordinates of the vertices: on the benchmark lattice faces TOUCH support
// spheres exactly, and membership must not depend on how the closest point is computed.
// Mirrored quirk: `is_tiny(area())` / `is_tiny(norm2(ba))` compare dimensional quantities
// with the absolute tiny = 6e-6, so a fine wall mesh (triangle area below 6e-6, e.g. C5) is
// clamped to its longest EDGE and a short segment to its first vertex.
template<int DD> TIT_HD Vec<DD> clamp_to_segment(const Vec<DD>& a, const Vec<DD>& b, const Vec<DD>& p, double tiny) {
  const Vec<DD> ba = xsubv(b, a);
  const double len2 = xdot(ba, ba);
  if (fabs(len2) <= tiny) return a;
  const double t = xdot(xsubv(p, a), ba) / len2;
  if (t < 0.0) return a;
  if (t > 1.0) return b;
  Vec<DD> q;
  for (int d = 0; d < DD; ++d) q[d] = xadd(a[d], xmul(t, ba[d]));
  return q;
}
TIT_HD bool face_intersects(const FaceFrame<2>& f, const Vec<2>& c, double radius, double radius2, double tiny) {
  for (int d = 0; d < 2; ++d)
    if (!(c[d] - radius <= f.hi[d] && f.lo[d] <= c[d] + radius)) return false;
  Vec<2> a, b;
  for (int d = 0; d < 2; ++d) { a[d] = f.a[d]; b[d] = f.b[d]; }
  const Vec<2> x = xsubv(clamp_to_segment<2>(a, b, c, tiny), c);
  return xdot(x, x) <= radius2;
}

// 0: a proper triangle; 1 / 2 / 3: is_tiny(area) and the longest edge is ab / bc / ac (triangle.hpp:118-135).
TIT_HD int triangle_degeneracy(const Vec<3>& a, const Vec<3>& b, const Vec<3>& c, double tiny) {
  const Vec<3> ba = xsubv(b, a), ca = xsubv(c, a), cb = xsubv(c, b);
  Vec<3> w;
  w[0] = xmul(xsub(xmul(ba[1], ca[2]), xmul(ba[2], ca[1])), 0.5);
  w[1] = xmul(xsub(xmul(ba[2], ca[0]), xmul(ba[0], ca[2])), 0.5);
  w[2] = xmul(xsub(xmul(ba[0], ca[1]), xmul(ba[1], ca[0])), 0.5);
  if (!(fabs(sqrt(xdot(w, w))) <= tiny)) return 0;
  const double ab = xdot(ba, ba), bc = xdot(cb, cb), ca2 = xdot(ca, ca);
  if (ab >= bc && ab >= ca2) return 1;
  if (bc >= ca2) return 2;
  return 3;
}
// The 80 bytes of a triangle the exact test reads (the face search gathers these per candidate
// face; the ~350-byte frame is only needed by the edge integrals).
struct FaceGeom3 {
  double a[3], b[3], c[3];
  int degen, pad;
};
static_assert(sizeof(FaceGeom3) == 80, "FaceGeom3 must stay 80 bytes");
TIT_HD bool face_intersects3(const FaceGeom3& f, const Vec<3>& p, double radius, double radius2, double tiny) {
  Vec<3> a, b, c, q;
  for (int d = 0; d < 3; ++d) { a[d] = f.a[d]; b[d] = f.b[d]; c[d] = f.c[d]; }
  for (int d = 0; d < 3; ++d) {  // bbox of the triangle (geom/triangle.hpp box(): min / max of the vertices)
    const double lo = fmin(fmin(a[d], b[d]), c[d]), hi = fmax(fmax(a[d], b[d]), c[d]);
    if (!(p[d] - radius <= hi && lo <= p[d] + radius)) return false;
  }
  if (f.degen) {
    q = f.degen == 1 ? clamp_to_segment<3>(a, b, p, tiny) : f.degen == 2 ? clamp_to_segment<3>(b, c, p, tiny) : clamp_to_segment<3>(a, c, p, tiny);
  } else {
    // Ericson's region walk, geom/triangle.hpp:137-181.
    const Vec<3> ba = xsubv(b, a), ca = xsubv(c, a);
    const Vec<3> pa = xsubv(p, a);
    const double d1 = xdot(ba, pa), d2 = xdot(ca, pa);
    const Vec<3> pb = xsubv(p, b);
    const double d3 = xdot(ba, pb), d4 = xdot(ca, pb);
    const Vec<3> pc = xsubv(p, c);
    const double d5 = xdot(ba, pc), d6 = xdot(ca, pc);
    const double vc = xsub(xmul(d1, d4), xmul(d3, d2));
    const double vb = xsub(xmul(d5, d2), xmul(d1, d6));
    const double va = xsub(xmul(d3, d6), xmul(d5, d4));
    if (d1 <= 0.0 && d2 <= 0.0) q = a;
    else if (d3 >= 0.0 && d4 <= d3) q = b;
    else if (d6 >= 0.0 && d5 <= d6) q = c;
    else if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
      const double t = d1 / xsub(d1, d3);
      for (int d = 0; d < 3; ++d) q[d] = xadd(a[d], xmul(t, ba[d]));
    } else if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
      const double t = d2 / xsub(d2, d6);
      for (int d = 0; d < 3; ++d) q[d] = xadd(a[d], xmul(t, ca[d]));
    } else if (va <= 0.0 && d4 >= d3 && d5 >= d6) {
      const double t = xsub(d4, d3) / xadd(xsub(d4, d3), xsub(d5, d6));
      const Vec<3> cb = xsubv(c, b);
      for (int d = 0; d < 3; ++d) q[d] = xadd(b[d], xmul(t, cb[d]));
    } else {
      const double den = xadd(xadd(va, vb), vc);
      const double v = vb / den, w = vc / den;
      for (int d = 0; d < 3; ++d) q[d] = xadd(xadd(a[d], xmul(v, ba[d])), xmul(w, ca[d]));
    }
  }
  const Vec<3> x = xsubv(q, p);
  return xdot(x, x) <= radius2;
}
}  // namespace titgpu
