// Slab decomposition across GPUs: device side of the ghost-layer exchange.
//
// Replaces, across GPUs, what the reference's block partition does across threads
// (/root/reference/source/tit/sph/particle_mesh.hpp:165-241, geom/partition/*): a rank
// owns the fluid particles of one slab [lo, hi) along `axis` and keeps GHOST copies of the
// neighbouring slabs' particles within `halo` of its slab; ghosts are neighbours only.
//
// Rank-local particle ids ("orig"): [0, n_owned) owned fluid, [n_owned, nf) ghosts (those
// of the left neighbour first), [nf, n) the rank's wall particles. Per step
// (time_integrator.hpp:161-184):
//   begin_step   particles that left the slab migrate to the neighbour; then every rank
//                selects the owned particles within `halo` of a neighbour (the halo carries a
//                margin for the motion within the step, so the SET stays valid until the
//                next begin_step), sends their records and rebuilds its arrays as
//                owned | ghosts | walls. Counts travel first (two small host exchanges);
//   refresh      before the later neighbour searches of the step the same set is re-sent
//                with its current records (fixed sizes: no host synchronisation);
//   N / phi      after the shifting sums (fluid_equations.hpp:337-426) the owners publish
//                {N, phi} of the set: the near-surface pass reads them from neighbours
//                (:440-452);
//   shifted      after apply_shifts the owners publish the shifted records: the free-surface
//                correction reads neighbours' shifted positions and densities (:489-511);
//   dt           MIN / MAX all-reduce of the time-step scalars (:203-221).
// Every list is built by prefix sums in local-id order: the ghost numbering, hence the
// order of every floating-point sum, does not depend on thread scheduling.
#pragma once

#include "mg_transport.h"

namespace titgpu {

template<int D> __device__ __forceinline__ double mg_axis_coord(const double4& a, int axis) {
  if (axis == 0) return a.x;
  if (axis == 1) return a.y;
  return D == 3 ? a.z : a.y;
}

// Owned particles that left the slab; also the inverse of the sort permutation.
template<int D>
__global__ void k_mg_classify(const double4* __restrict__ A, const int* __restrict__ orig, int n, int n_owned, int axis, double lo, double hi, int has_l, int has_r,
                              int* __restrict__ pos_of, int* __restrict__ keep, int* __restrict__ go_l, int* __restrict__ go_r) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int o = orig[a];
  pos_of[o] = a;
  if (o >= n_owned) return;
  const double x = mg_axis_coord<D>(A[a], axis);
  const int st = (has_l && x < lo) ? 1 : (has_r && x >= hi) ? 2 : 0;
  keep[o] = st == 0;
  go_l[o] = st == 1;
  go_r[o] = st == 2;
}

static __global__ void k_mg_scatter_owned(const double4* __restrict__ A, const double4* __restrict__ B, const long long* __restrict__ gid, const int* __restrict__ pos_of, int n_owned,
                                          const int* __restrict__ keep, const int* __restrict__ go_l, const int* __restrict__ s_keep, const int* __restrict__ s_l, const int* __restrict__ s_r,
                                          double4* __restrict__ A_o, double4* __restrict__ B_o, long long* __restrict__ gid_o, double4* __restrict__ mA_l, double4* __restrict__ mB_l,
                                          long long* __restrict__ mg_l, double4* __restrict__ mA_r, double4* __restrict__ mB_r, long long* __restrict__ mg_r) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_owned) return;
  const int a = pos_of[o];
  const double4 ra = A[a], rb = B[a];
  const long long g = gid[o];
  if (keep[o]) { const int k = s_keep[o]; A_o[k] = ra; B_o[k] = rb; gid_o[k] = g; }
  else if (go_l[o]) { const int k = s_l[o]; mA_l[k] = ra; mB_l[k] = rb; mg_l[k] = g; }
  else { const int k = s_r[o]; mA_r[k] = ra; mB_r[k] = rb; mg_r[k] = g; }
}

// Which owned particles (canonical order, after the migration) the neighbouring slab needs as
// ghosts: everything within `halo_pair` of it (the neighbours of its own particles: one support
// radius + margin), and up to `halo` (two radii + the wall-face edge + margin) the particles that
// can lie within a support radius of a WALL particle - the neighbour extrapolates the density of
// its wall particles from them (fluid_equations.hpp:122-164); away from the walls nothing beyond
// one radius is ever read. (fflag / CF_WALL: a boundary face within reach of the particle's
// face-grid cell.) A particle outside its owner's slab at this point crossed more than one slab
// within a step: flagged, the host reports it.
template<int D>
__global__ void k_mg_band(const double4* __restrict__ A, int n_owned, int axis, double lo, double hi, double halo_pair, double halo, GridDesc fg, const unsigned char* __restrict__ fflag, int has_l,
                          int has_r, int* __restrict__ in_l, int* __restrict__ in_r, int* __restrict__ bad) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_owned) return;
  const double4 a = A[o];
  const double x = mg_axis_coord<D>(a, axis);
  const bool l2 = has_l && x < lo + halo, r2 = has_r && x >= hi - halo;
  bool near_wall = true;
  if ((l2 && !(x < lo + halo_pair)) || (r2 && !(x >= hi - halo_pair))) {
    Vec<D> r;
    r[0] = a.x; r[1] = a.y;
    if constexpr (D == 3) r[2] = a.z;
    int fci[D];
    cell_coords<D>(fg, r, fci);
    near_wall = (fflag[cell_flat<D>(fg, fci)] & CF_WALL) != 0;
  }
  in_l[o] = l2 && (x < lo + halo_pair || near_wall);
  in_r[o] = r2 && (x >= hi - halo_pair || near_wall);
  if ((has_l && x < lo) || (has_r && x >= hi)) *bad = 1;
}

static __global__ void k_mg_pack_halo(const double4* __restrict__ A, const double4* __restrict__ B, int n_owned, const int* __restrict__ in_l, const int* __restrict__ in_r,
                                      const int* __restrict__ s_l, const int* __restrict__ s_r, int off_r, double4* __restrict__ sA, double4* __restrict__ sB, int* __restrict__ idx) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_owned) return;
  const bool l = in_l[o] != 0, r = in_r[o] != 0;
  if (!l && !r) return;
  const double4 ra = A[o], rb = B[o];
  if (l) { const int k = s_l[o]; sA[k] = ra; sB[k] = rb; idx[k] = o; }
  if (r) { const int k = off_r + s_r[o]; sA[k] = ra; sB[k] = rb; idx[k] = o; }
}

static __global__ void k_mg_inverse(const int* __restrict__ orig, int n, int* __restrict__ pos_of) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n) pos_of[orig[a]] = a;
}
// dst[k] = src[position of the k-th member of the halo set]
static __global__ void k_mg_gather4(const double4* __restrict__ src, const int* __restrict__ pos_of, const int* __restrict__ idx, int cnt, double4* __restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) dst[k] = src[pos_of[idx[k]]];
}
// dst[position of ghost k] = src[k]
static __global__ void k_mg_scatter4(const double4* __restrict__ src, const int* __restrict__ pos_of, int n_owned, int cnt, double4* __restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) dst[pos_of[n_owned + k]] = src[k];
}
template<int D>
__global__ void k_mg_gather_nphi(const double* __restrict__ N_s, const double* __restrict__ phi_s, const int* __restrict__ pos_of, const int* __restrict__ idx, int cnt, double4* __restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int a = pos_of[idx[k]];
  const Vec<D> N = load_vec<D>(N_s, a);
  dst[k] = make_double4(N[0], N[1], D == 3 ? N[D - 1] : 0.0, phi_s[a]);
}
template<int D>
__global__ void k_mg_scatter_nphi(const double4* __restrict__ src, const int* __restrict__ pos_of, int n_owned, int cnt, GridDesc g, const double4* __restrict__ A, double* __restrict__ N_s,
                                  double* __restrict__ phi_s, unsigned char* __restrict__ fs_flag, unsigned char* __restrict__ cell_fs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int a = pos_of[n_owned + k];
  const double4 v = src[k];
  Vec<D> N;
  N[0] = v.x; N[1] = v.y;
  if constexpr (D == 3) N[2] = v.z;
  store_vec<D>(N_s, a, N);
  phi_s[a] = v.w;
  const bool fs = bits_equal(v.w, kPhiMin);
  fs_flag[a] = fs ? 1 : 0;
  if (fs) {
    Vec<D> r;
    double rho_unused;
    Pack<D>::pos(A, a, r, rho_unused);
    int ci[D];
    cell_coords<D>(g, r, ci);
    thread_mark_cell_block<D>(g, ci, cell_fs);  // as the owner's warp does for its own particles (shift_finish)
  }
}

// Owned records in local-id order (download / checks).
static __global__ void k_mg_gather_owned(const double4* __restrict__ A, const double4* __restrict__ B, const int* __restrict__ orig, int n, int n_owned, double4* __restrict__ oA, double4* __restrict__ oB) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int o = orig[a];
  if (o >= n_owned) return;
  oA[o] = A[a];
  oB[o] = B[a];
}
static __global__ void k_mg_iota64(long long* __restrict__ p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

}  // namespace titgpu
