// Small device/host vector algebra and shared parameter blocks for the
// B200 WCSPH step. Semantics (tolerances, summation order of `dot`, LU without
// pivoting) follow the reference's core numerics:
//   /root/reference/source/tit/core/math.hpp:145-173   tiny, is_tiny, bitwise_equal
//   /root/reference/source/tit/core/_vec/vec.hpp:640-700 dot/norm/normalize/cross
//   /root/reference/source/tit/core/_mat/fact.hpp:84-108 lu
#pragma once

#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdint>

#ifndef TIT_HD
#define TIT_HD __host__ __device__ __forceinline__
#endif
#ifndef TIT_HDN
#define TIT_HDN __host__ __device__ __noinline__
#endif

namespace titgpu {

template<int D>
struct Vec {
  double e[D];
  TIT_HD double& operator[](int i) { return e[i]; }
  TIT_HD const double& operator[](int i) const { return e[i]; }
};
template<int D>
struct Mat {
  Vec<D> r[D];
  TIT_HD Vec<D>& operator[](int i) { return r[i]; }
  TIT_HD const Vec<D>& operator[](int i) const { return r[i]; }
};

template<int D> TIT_HD Vec<D> vzero() { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = 0.0; return r; }
template<int D> TIT_HD Mat<D> mzero() { Mat<D> r; for (int i = 0; i < D; ++i) r[i] = vzero<D>(); return r; }
template<int D> TIT_HD Mat<D> meye() { Mat<D> r = mzero<D>(); for (int i = 0; i < D; ++i) r[i][i] = 1.0; return r; }
template<int D> TIT_HD Vec<D> operator+(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] + b[i]; return r; }
template<int D> TIT_HD Vec<D> operator-(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] - b[i]; return r; }
template<int D> TIT_HD Vec<D> operator*(double s, const Vec<D>& a) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = s * a[i]; return r; }
template<int D> TIT_HD Vec<D> operator*(const Vec<D>& a, double s) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = a[i] * s; return r; }
template<int D> TIT_HD Vec<D>& operator+=(Vec<D>& a, const Vec<D>& b) { for (int i = 0; i < D; ++i) a[i] += b[i]; return a; }
template<int D> TIT_HD Vec<D>& operator-=(Vec<D>& a, const Vec<D>& b) { for (int i = 0; i < D; ++i) a[i] -= b[i]; return a; }
template<int D> TIT_HD double dot(const Vec<D>& a, const Vec<D>& b) { double r = a[0] * b[0]; for (int i = 1; i < D; ++i) r += a[i] * b[i]; return r; }
template<int D> TIT_HD double norm2(const Vec<D>& a) { return dot(a, a); }
template<int D> TIT_HD double norm(const Vec<D>& a) { return sqrt(norm2(a)); }
template<int D> TIT_HD Vec<D> normalize(const Vec<D>& a, double tiny2) {
  const double n2 = norm2(a);
  if (n2 >= tiny2) return a * (1.0 / sqrt(n2));
  return vzero<D>();
}
TIT_HD Vec<3> cross(const Vec<3>& a, const Vec<3>& b) {
  Vec<3> r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}

// Fused-free arithmetic for the discrete decisions that must agree bit for bit
// with the oracle: the inclusive neighbour predicate |r_a - r_b|^2 <= (2h)^2
// (geom/bsphere.hpp:52-53) and the containment winding number. nvcc contracts
// a*b+c into FMA by default; these intrinsics are never contracted.
#ifdef __CUDA_ARCH__
TIT_HD double xmul(double a, double b) { return __dmul_rn(a, b); }
TIT_HD double xadd(double a, double b) { return __dadd_rn(a, b); }
TIT_HD double xsub(double a, double b) { return __dsub_rn(a, b); }
#else
TIT_HD double xmul(double a, double b) { return a * b; }
TIT_HD double xadd(double a, double b) { return a + b; }
TIT_HD double xsub(double a, double b) { return a - b; }
#endif
template<int D> TIT_HD double xdot(const Vec<D>& a, const Vec<D>& b) {
  double r = xmul(a[0], b[0]);
  for (int i = 1; i < D; ++i) r = xadd(r, xmul(a[i], b[i]));
  return r;
}
template<int D> TIT_HD Vec<D> xsubv(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; ++i) r[i] = xsub(a[i], b[i]); return r; }

template<int D> TIT_HD Vec<D> load_vec(const double* p, size_t i) { Vec<D> r; for (int d = 0; d < D; ++d) r[d] = p[i * D + d]; return r; }
template<int D> TIT_HD void store_vec(double* p, size_t i, const Vec<D>& v) { for (int d = 0; d < D; ++d) p[i * D + d] = v[d]; }
template<int D> TIT_HD void store_mat(double* p, size_t i, const Mat<D>& m) { for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) p[(i * D + a) * D + b] = m[a][b]; }

template<int D> TIT_HD Mat<D> transpose(const Mat<D>& A) { Mat<D> R; for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) R[i][j] = A[j][i]; return R; }
// (A b)_k = sum_i A[k][i] b[i], i ascending (core/_mat/mat.hpp:147-152).
template<int D> TIT_HD Vec<D> matvec(const Mat<D>& A, const Vec<D>& b) {
  Vec<D> r;
  for (int k = 0; k < D; ++k) { double s = A[k][0] * b[0]; for (int i = 1; i < D; ++i) s += A[k][i] * b[i]; r[k] = s; }
  return r;
}
template<int D> TIT_HD Mat<D> matmul(const Mat<D>& A, const Mat<D>& B) {
  Mat<D> R = mzero<D>();
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) for (int k = 0; k < D; ++k) R[i][j] += A[i][k] * B[k][j];
  return R;
}
// LU without pivoting; fails on a tiny pivot (core/_mat/fact.hpp:84-108), then
// inverse via unit-lower / upper solves of the identity (part.hpp:100-125).
template<int D> TIT_HD bool lu_inverse(const Mat<D>& A, Mat<D>& inv, double tiny) {
  // Every loop has compile-time bounds after unrolling: the factors stay in
  // registers (dynamically indexed copies would live in local memory).
  Mat<D> LU = mzero<D>();
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (j < i) {
        double s = A[i][j];
#pragma unroll
        for (int k = 0; k < D; ++k) if (k < j) s -= LU[i][k] * LU[k][j];
        LU[i][j] = s / LU[j][j];
      }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (j >= i) {
        double s = A[i][j];
#pragma unroll
        for (int k = 0; k < D; ++k) if (k < i) s -= LU[i][k] * LU[k][j];
        LU[i][j] = s;
      }
    }
    if (fabs(LU[i][i]) <= tiny) return false;
  }
  Mat<D> x = meye<D>();
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) if (j < i) x[i] -= LU[i][j] * x[j];
  }
#pragma unroll
  for (int ii = 0; ii < D; ++ii) {
    const int i = D - 1 - ii;
#pragma unroll
    for (int j = 0; j < D; ++j) if (j > i) x[i] -= LU[i][j] * x[j];
    const double dinv = LU[i][i];
#pragma unroll
    for (int c = 0; c < D; ++c) x[i][c] = x[i][c] / dinv;
  }
  inv = x;
  return true;
}

TIT_HD bool bits_equal(double a, double b) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(a) == __double_as_longlong(b);
#else
  union { double d; long long l; } x{a}, y{b};
  return x.l == y.l;
#endif
}

// Constants of fluid_equations.hpp:518-524.
constexpr double kCFL = 0.4;
constexpr double kCForce = 0.25;
constexpr double kCVisc = 0.125;
constexpr double kCShift = 0.2;
constexpr double kPhiMax = 1.0;
constexpr double kPhiMin = DBL_MIN;

// Particle cells per support radius: a particle's neighbours lie within +-KC_
// cells per axis.
constexpr int KC_ = 2;

// Uniform cell grid. Two instances share one origin: the particle hash (cells
// of radius / KC_) and the static face index (cells of one radius). The face
// cell size is radius * (1 + 2^-20) so that points exactly one radius apart
// (the initial lattice) are always within the expected cell distance despite
// rounding of the quotient.
struct GridDesc {
  double org[3];
  double cinv;  // 1 / cell size
  int nc[3];
  int ncells;
  int tyl;  // 3-D search grid: log2 of the tile width along y of the cell order (0 = plain row-major), see col_base
};

struct Params {
  double g, mu, cs0, rho0, xi, h;
  double hinv, radius, radius2, tiny, tiny2;
  double w_val;   // weight<D> h^-D
  double w_flux;  // weight<D> h^-1
  double w_anti;  // weight<D>
  double inv_rho0, tait_b;  // 1 / rho0 and rho0 cs0^2 / xi: loop invariants of the pair passes, read from the constant bank
  double k_fs;    // -log(0.05) / 0.01^2
  double cos_fov2;  // cos(pi/4)^2 as evaluated in double (fluid_equations.hpp:406-407)
  int eos;
  int nf, nx, n;
  int n_owned;    // slab decomposition: fluid particles [n_owned, nf) are ghosts (neighbours only); == nf otherwise
  float pre_thr;  // FP32 pre-filter threshold on the squared distance in cell units
  float face_thr; // same for the face cull (face-grid cell units)
  float list_thr; // FP32 threshold of the candidate-list build: ((radius + skin) / cell)^2 + margin
  double skin_half2;  // (skin / 2)^2: a particle that moves farther than this from where the lists were built invalidates them
  float oor;      // |grid coordinate| beyond which the FP32 pre-filter is bypassed
  GridDesc grid;   // particle hash
  GridDesc fgrid;  // face index + per-cell wall / containment flags
};

// x^e: small integer exponents (the Tait defaults xi = 7 -> 7, 6, 3) by repeated
// multiplication, anything else through pow().
TIT_HD double pow_fast(double x, double e) {
  if (e == 7.0) { const double x2 = x * x, x4 = x2 * x2; return x4 * x2 * x; }
  if (e == 6.0) { const double x2 = x * x; return x2 * x2 * x2; }
  if (e == 3.0) return x * x * x;
  return pow(x, e);
}

// Equation of state (sph/equation_of_state.hpp:19-122).
struct Eos {
  TIT_HD static double p(const Params& P, double rho) {
    if (P.eos == 1) return P.cs0 * P.cs0 * (rho - P.rho0);
    const double B = P.rho0 * (P.cs0 * P.cs0) / P.xi;
    return B * (pow_fast(rho / P.rho0, P.xi) - 1.0);
  }
  TIT_HD static double cs(const Params& P, double rho) {
    if (P.eos == 1) return P.cs0;
    return P.cs0 * pow_fast(rho / P.rho0, (P.xi - 1.0) / 2.0);
  }
  TIT_HD static double H(const Params& P, double rho) {
    if (P.eos == 1) return P.cs0 * P.cs0 * log(rho / P.rho0);
    const double x1 = P.xi - 1.0;
    return P.cs0 * P.cs0 * (pow_fast(rho / P.rho0, x1) - 1.0) / x1;
  }
  TIT_HD static double rho_from_H(const Params& P, double Hh) {
    if (P.eos == 1) return P.rho0 * exp(Hh / (P.cs0 * P.cs0));
    const double x1 = P.xi - 1.0;
    return P.rho0 * pow(1.0 + x1 * Hh / (P.cs0 * P.cs0), 1.0 / x1);
  }
};

}  // namespace titgpu
