// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// CPU restatement of the reference's small-vector numerics and geometry for the
// WCSPH hot path. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Follows (all under /root/reference/source/tit/):
//   core/math.hpp:145-173          tiny_v, is_tiny, approx_equal_to, bitwise_equal
//   core/_vec/vec.hpp:640-720      dot, norm2, norm, normalize, cross, det
//   core/_mat/fact.hpp:84-108      lu (Doolittle, no pivoting, tiny-pivot failure)
//   core/_mat/part.hpp:100-125     triangular solves
//   core/_mat/mat.hpp:147-162      Mat*Vec, Mat*Mat
//   geom/bbox.hpp, geom/bsphere.hpp:52-53, geom/grid.hpp:90-168
//   geom/segment.hpp:61-112, geom/triangle.hpp:78-186
//   geom/winding/exact_winding.hpp:32-43
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

// core/math.hpp:145-146: cbrt(eps) ~ 6.0554544523933e-6.
inline const double tiny = std::pow(std::numeric_limits<double>::epsilon(), 1.0 / 3.0);

inline bool is_tiny(double a) { return std::abs(a) <= tiny; }
inline bool is_tiny(double a, double eps) { return std::abs(a) <= eps; }
inline bool approx_equal(double a, double b) { return is_tiny(a - b); }
inline bool bitwise_equal(double a, double b) {
  std::uint64_t x, y;
  std::memcpy(&x, &a, 8);
  std::memcpy(&y, &b, 8);
  return x == y;
}
inline double pow2(double a) { return a * a; }

// Packed small vector / row-major matrix (own types so that `int D` deduces).
template<int D>
struct Vec {
  double e[D];
  double& operator[](std::size_t i) { return e[i]; }
  const double& operator[](std::size_t i) const { return e[i]; }
  void fill(double x) { for (int i = 0; i < D; ++i) e[i] = x; }
  double* data() { return e; }
  const double* data() const { return e; }
};
template<int D>
struct Mat {
  Vec<D> rows[D];
  Vec<D>& operator[](std::size_t i) { return rows[i]; }
  const Vec<D>& operator[](std::size_t i) const { return rows[i]; }
  double* data() { return rows[0].e; }
};

template<int D> inline Vec<D> operator+(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] + b[i]; return r; }
template<int D> inline Vec<D> operator-(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] - b[i]; return r; }
template<int D> inline Vec<D> operator-(const Vec<D>& a) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = -a[i]; return r; }
template<int D> inline Vec<D> operator*(double s, const Vec<D>& a) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = s * a[i]; return r; }
template<int D> inline Vec<D> operator*(const Vec<D>& a, double s) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] * s; return r; }
template<int D> inline Vec<D> operator/(const Vec<D>& a, double s) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] / s; return r; }
template<int D> inline Vec<D>& operator+=(Vec<D>& a, const Vec<D>& b) { for (int i = 0; i < D; ++i) a[i] += b[i]; return a; }
template<int D> inline Vec<D>& operator-=(Vec<D>& a, const Vec<D>& b) { for (int i = 0; i < D; ++i) a[i] -= b[i]; return a; }

// core/_vec/vec.hpp:644-646: left-to-right a0*b0 + a1*b1 (+ a2*b2).
template<int D> inline double dot(const Vec<D>& a, const Vec<D>& b) {
  double r = a[0] * b[0];
  for (int i = 1; i < D; ++i) r += a[i] * b[i];
  return r;
}
template<int D> inline double norm2(const Vec<D>& a) { return dot(a, a); }
template<int D> inline double norm(const Vec<D>& a) { return std::sqrt(norm2(a)); }
// core/_vec/vec.hpp:663-678: zero for vectors shorter than tiny.
template<int D> inline Vec<D> normalize(const Vec<D>& a) {
  const double n2 = norm2(a);
  if (n2 >= tiny * tiny) return a / std::sqrt(n2);
  return Vec<D>{};
}
inline Vec<2> cross(const Vec<2>& a) { return {a[1], -a[0]}; }
inline Vec<3> cross(const Vec<3>& a, const Vec<3>& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline double det(const Vec<2>& a, const Vec<2>& b) { return dot(a, cross(b)); }
inline double det(const Vec<3>& a, const Vec<3>& b, const Vec<3>& c) { return dot(a, cross(b, c)); }

template<int D> inline Mat<D> outer(const Vec<D>& a, const Vec<D>& b) {
  Mat<D> R;
  for (int i = 0; i < D; ++i) R[i] = a[i] * b;
  return R;
}
template<int D> inline Mat<D> transpose(const Mat<D>& A) {
  Mat<D> R;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) R[i][j] = A[j][i];
  return R;
}
template<int D> inline Mat<D> eye() {
  Mat<D> R{};
  for (int i = 0; i < D; ++i) R[i][i] = 1.0;
  return R;
}
// core/_mat/mat.hpp:147-152.
template<int D> inline Vec<D> matvec(const Mat<D>& A, const Vec<D>& b) {
  const Mat<D> T = transpose(A);
  Vec<D> r = T[0] * b[0];
  for (int i = 1; i < D; ++i) r += T[i] * b[i];
  return r;
}
// core/_mat/mat.hpp:155-163.
template<int D> inline Mat<D> matmul(const Mat<D>& A, const Mat<D>& B) {
  Mat<D> R{};
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j)
      for (int k = 0; k < D; ++k) R[i][j] += A[i][k] * B[k][j];
  return R;
}
template<int D> inline Mat<D>& operator+=(Mat<D>& A, const Mat<D>& B) { for (int i = 0; i < D; ++i) A[i] += B[i]; return A; }
template<int D> inline Mat<D>& operator-=(Mat<D>& A, const Mat<D>& B) { for (int i = 0; i < D; ++i) A[i] -= B[i]; return A; }
template<int D> inline Mat<D> operator*(double s, const Mat<D>& A) { Mat<D> R; for (int i = 0; i < D; ++i) R[i] = s * A[i]; return R; }
// core/_mat/mat.hpp:166-169: division multiplies by the reciprocal.
template<int D> inline Mat<D> operator/(const Mat<D>& A, double s) { const double inv = 1.0 / s; Mat<D> R; for (int i = 0; i < D; ++i) R[i] = A[i] * inv; return R; }

// LU factorisation + inverse (core/_mat/fact.hpp:60-108, part.hpp:100-125).
template<int D> inline bool lu_inverse(const Mat<D>& A, Mat<D>& inv) {
  Mat<D> LU{};
  for (int i = 0; i < D; ++i) {
    for (int j = 0; j < i; ++j) {
      LU[i][j] = A[i][j];
      for (int k = 0; k < j; ++k) LU[i][j] -= LU[i][k] * LU[k][j];
      LU[i][j] /= LU[j][j];
    }
    for (int j = i; j < D; ++j) {
      LU[i][j] = A[i][j];
      for (int k = 0; k < i; ++k) LU[i][j] -= LU[i][k] * LU[k][j];
    }
    if (is_tiny(LU[i][i])) return false;
  }
  // solve(eye): x rows are Vec (matrix right-hand side, row i = x[i]).
  Mat<D> x = eye<D>();
  for (int i = 0; i < D; ++i) {  // lower_unit
    for (int j = 0; j < i; ++j) x[i] -= LU[i][j] * x[j];
    x[i] = x[i] / 1.0;
  }
  for (int i = D - 1; i >= 0; --i) {  // upper_diag
    for (int j = i + 1; j < D; ++j) x[i] -= LU[i][j] * x[j];
    x[i] = x[i] / LU[i][i];
  }
  inv = x;
  return true;
}

// ---- geom ---------------------------------------------------------------

template<int D> struct BBox {
  Vec<D> lo{}, hi{};
  BBox() = default;
  explicit BBox(const Vec<D>& p) : lo(p), hi(p) {}
  Vec<D> extents() const { return hi - lo; }
  BBox& grow(const Vec<D>& a) { lo -= a; hi += a; return *this; }
  BBox& grow(double a) { Vec<D> v; v.fill(a); return grow(v); }
  BBox& shrink(const Vec<D>& a) { lo += a; hi -= a; return *this; }
  BBox& expand(const Vec<D>& p) { for (int i = 0; i < D; ++i) { lo[i] = std::min(lo[i], p[i]); hi[i] = std::max(hi[i], p[i]); } return *this; }
  BBox& intersect(const BBox& b) { for (int i = 0; i < D; ++i) { lo[i] = std::max(lo[i], b.lo[i]); hi[i] = std::min(hi[i], b.hi[i]); } return *this; }
  bool intersects(const BBox& b) const { for (int i = 0; i < D; ++i) if (!(lo[i] <= b.hi[i] && b.lo[i] <= hi[i])) return false; return true; }
};

// geom/bsphere.hpp:47-53 (inclusive).
template<int D> struct BSphere {
  Vec<D> c; double r;
  BBox<D> box() const { return BBox<D>(c).grow(r); }
  bool contains(const Vec<D>& p) const { return norm2(p - c) <= pow2(r); }
};

// geom/grid.hpp:90-168.
template<int D> struct Grid {
  BBox<D> box;
  std::array<std::size_t, D> num{};
  Vec<D> ext{}, inv_ext{};
  void set_cell_extents(double hint) {
    const Vec<D> e = box.extents();
    for (int i = 0; i < D; ++i) {
      const double nf = std::max(std::ceil(e[i] / hint), 1.0);
      num[i] = static_cast<std::size_t>(nf);
      ext[i] = e[i] / nf;
      inv_ext[i] = 1.0 / ext[i];
    }
  }
  std::size_t flat_num() const { std::size_t p = 1; for (int i = 0; i < D; ++i) p *= num[i]; return p; }
  std::array<std::size_t, D> cell_index(const Vec<D>& p) const {
    std::array<std::size_t, D> r;
    for (int i = 0; i < D; ++i) r[i] = static_cast<std::size_t>((p[i] - box.lo[i]) * inv_ext[i]);
    return r;
  }
  std::size_t flatten(const std::array<std::size_t, D>& idx) const {
    std::size_t f = idx[0];
    for (int i = 1; i < D; ++i) f = num[i] * f + idx[i];
    return f;
  }
  // low/high inclusive cell range overlapping the search box. Returns false
  // when the intersection is empty (the reference leaves that case undefined).
  bool cells_intersecting(const BBox<D>& search, std::array<std::size_t, D>& lo, std::array<std::size_t, D>& hi) const {
    const Vec<D> half = ext / 2.0;
    BBox<D> s = search;
    s.grow(half).intersect(box).shrink(half);
    // The reference requires a non-empty intersection (grid.hpp:160) and has no
    // emptiness test. A search box that is truly disjoint from the grid box ends
    // up with lo - hi >= half a cell here; anything less is an overlap whose
    // lo/hi may cross by rounding only (e.g. a wall face lying exactly on the
    // upper side of the grid box) and must NOT be dropped.
    for (int i = 0; i < D; ++i)
      if (!(s.lo[i] - s.hi[i] <= half[i] * (1.0 - 1e-9))) return false;
    for (int i = 0; i < D; ++i) { s.lo[i] = std::max(s.lo[i], box.lo[i]); s.hi[i] = std::max(s.hi[i], box.lo[i]); }
    lo = cell_index(s.lo);
    hi = cell_index(s.hi);
    for (int i = 0; i < D; ++i) { lo[i] = std::min(lo[i], num[i] - 1); hi[i] = std::max(std::min(hi[i], num[i] - 1), lo[i]); }
    return true;
  }
};

// geom/search/kd_tree_search.hpp:33-142: the alternative search index (same contract
// as the grid index). Median split along the longest axis of the bounding box of the
// node's points (bipartition.hpp:101-125: nth_element on that coordinate; the median
// point stays in the node), recursive sphere query that always descends to the side
// of the centre and to the other side only if the splitting plane is strictly closer
// than the radius (:113-121). The reference allocates nodes from parallel tasks, so
// its node numbering is not defined; this one numbers them in depth-first order.
template<int D> struct KDTree {
  static constexpr std::size_t npos = std::numeric_limits<std::size_t>::max();
  struct Node { std::size_t index = 0, axis = 0, left = npos, right = npos; };
  const std::vector<Vec<D>>* pts = nullptr;
  std::vector<Node> nodes;

  explicit KDTree(const std::vector<Vec<D>>& points) : pts(&points) {
    if (points.empty()) return;
    std::vector<std::size_t> perm(points.size());
    for (std::size_t i = 0; i < perm.size(); ++i) perm[i] = i;
    nodes.reserve(points.size());
    build(perm.data(), perm.size());
  }
  std::size_t build(std::size_t* perm, std::size_t n) {
    const std::size_t id = nodes.size();
    nodes.emplace_back();
    if (n == 1) { nodes[id].index = perm[0]; return id; }
    BBox<D> box((*pts)[perm[0]]);
    for (std::size_t k = 1; k < n; ++k) box.expand((*pts)[perm[k]]);
    const Vec<D> e = box.extents();
    std::size_t axis = 0;  // max_value_index: the first of equal maxima
    for (int i = 1; i < D; ++i) if (e[i] > e[axis]) axis = i;
    const std::size_t med = n / 2;
    std::nth_element(perm, perm + med, perm + n, [&](std::size_t a, std::size_t b) { return (*pts)[a][axis] < (*pts)[b][axis]; });
    nodes[id].axis = axis;
    nodes[id].index = perm[med];
    if (med > 0) { const std::size_t l = build(perm, med); nodes[id].left = l; }
    if (n - med > 1) { const std::size_t r = build(perm + med + 1, n - med - 1); nodes[id].right = r; }
    return id;
  }
  template<class Out> void search(const BSphere<D>& s, Out&& out, std::size_t node = 0) const {
    if (node == npos || nodes.empty()) return;
    const Node& nd = nodes[node];
    const Vec<D>& p = (*pts)[nd.index];
    if (s.contains(p)) out(nd.index);
    const double delta = s.c[nd.axis] - p[nd.axis];
    if (delta < 0.0) {
      search(s, out, nd.left);
      if (pow2(delta) < pow2(s.r)) search(s, out, nd.right);
    } else {
      search(s, out, nd.right);
      if (pow2(delta) < pow2(s.r)) search(s, out, nd.left);
    }
  }
};

// geom/segment.hpp.
struct Segment {
  Vec<2> a, b;
  Vec<2> ba() const { return b - a; }
  BBox<2> box() const { return BBox<2>(a).expand(b); }
  Vec<2> center() const { return (a + b) / 2.0; }
  Vec<2> normal() const { return normalize(cross(ba())); }
  double length() const { return norm(cross(ba())); }
  double winding_number(const Vec<2>& p) const {
    const Vec<2> ap = a - p, bp = b - p;
    return std::atan2(det(ap, bp), dot(ap, bp)) / (2.0 * M_PI);
  }
  std::array<double, 2> project(const Vec<2>& o) const {
    const Vec<2> e = normalize(ba());
    return {dot(a - o, e), dot(b - o, e)};
  }
  Vec<2> clamp(const Vec<2>& p) const {
    const double len2 = norm2(ba());
    if (is_tiny(len2)) return a;
    const double t = dot(p - a, ba()) / len2;
    if (t < 0.0) return a;
    if (t > 1.0) return b;
    return a + t * ba();
  }
  bool intersects(const BSphere<2>& s) const { return s.box().intersects(box()) && s.contains(clamp(s.c)); }
};

// 3-D segment clamp used by the degenerate-triangle branch (triangle.hpp:118-135).
inline Vec<3> clamp_segment3(const Vec<3>& a, const Vec<3>& b, const Vec<3>& p) {
  const Vec<3> ba = b - a;
  const double len2 = norm2(ba);
  if (is_tiny(len2)) return a;
  const double t = dot(p - a, ba) / len2;
  if (t < 0.0) return a;
  if (t > 1.0) return b;
  return a + t * ba;
}

// geom/triangle.hpp.
struct Triangle {
  Vec<3> a, b, c;
  Vec<3> ba() const { return b - a; }
  Vec<3> cb() const { return c - b; }
  Vec<3> ca() const { return c - a; }
  BBox<3> box() const { return BBox<3>(a).expand(b).expand(c); }
  Vec<3> center() const { return (a + b + c) / 3.0; }
  Vec<3> wnormal() const { return cross(ba(), ca()) / 2.0; }
  Vec<3> normal() const { return normalize(wnormal()); }
  double area() const { return norm(wnormal()); }
  double winding_number(const Vec<3>& p) const {
    const Vec<3> ap = a - p, bp = b - p, cp = c - p;
    const double an = norm(ap), bn = norm(bp), cn = norm(cp);
    const double den = an * bn * cn + dot(ap, bp) * cn + dot(bp, cp) * an + dot(cp, ap) * bn;
    return std::atan2(det(ap, bp, cp), den) / (2.0 * M_PI);
  }
  std::array<Vec<2>, 3> project(const Vec<3>& o) const {
    const Vec<3> e1 = normalize(ba());
    const Vec<3> e2 = normalize(cross(wnormal(), e1));
    return {Vec<2>{dot(a - o, e1), dot(a - o, e2)}, Vec<2>{dot(b - o, e1), dot(b - o, e2)}, Vec<2>{dot(c - o, e1), dot(c - o, e2)}};
  }
  Vec<3> clamp(const Vec<3>& p) const {
    if (is_tiny(area())) {
      const double ab = norm2(ba()), bc = norm2(cb()), ca_ = norm2(ca());
      if (ab >= bc && ab >= ca_) return clamp_segment3(a, b, p);
      if (bc >= ca_) return clamp_segment3(b, c, p);
      return clamp_segment3(a, c, p);
    }
    const Vec<3> pa = p - a;
    const double d1 = dot(ba(), pa), d2 = dot(ca(), pa);
    if (d1 <= 0.0 && d2 <= 0.0) return a;
    const Vec<3> pb = p - b;
    const double d3 = dot(ba(), pb), d4 = dot(ca(), pb);
    if (d3 >= 0.0 && d4 <= d3) return b;
    const Vec<3> pc = p - c;
    const double d5 = dot(ba(), pc), d6 = dot(ca(), pc);
    if (d6 >= 0.0 && d5 <= d6) return c;
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) return a + (d1 / (d1 - d3)) * ba();
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) return a + (d2 / (d2 - d6)) * ca();
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && d4 >= d3 && d5 >= d6) return b + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * cb();
    const double v = vb / (va + vb + vc), w = vc / (va + vb + vc);
    return a + v * ba() + w * ca();
  }
  bool intersects(const BSphere<3>& s) const { return s.box().intersects(box()) && s.contains(clamp(s.c)); }
};

template<int D> struct FaceOf;
template<> struct FaceOf<2> { using type = Segment; };
template<> struct FaceOf<3> { using type = Triangle; };

// geom/surface.hpp (packed vertex array + D vertex indices per face).
template<int D> struct Surface {
  std::vector<Vec<D>> verts;
  std::vector<std::array<std::size_t, D>> faces;
  typename FaceOf<D>::type face(std::size_t f) const {
    if constexpr (D == 2) return Segment{verts[faces[f][0]], verts[faces[f][1]]};
    else return Triangle{verts[faces[f][0]], verts[faces[f][1]], verts[faces[f][2]]};
  }
  // geom/winding/exact_winding.hpp:32-43.
  bool contains(const Vec<D>& p) const {
    double w = 0.0;
    for (std::size_t f = 0; f < faces.size(); ++f) w += face(f).winding_number(p);
    return w > 0.5;
  }
};

}  // namespace orc
