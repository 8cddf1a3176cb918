// Shared-memory-staged pair passes (3-D, kernels of support radius 2h).
//
// One block = one TILE of 2 x 2 x 2 search cells (~64 particles on the lattice). Every
// neighbour of a tile particle lies in the 6 x 6 x 6 cells around the tile, and because the
// particles are sorted in row-major cell order those are 36 CONTIGUOUS runs of the record
// arrays (one per cell column, 6 cells long). The block
//   1. reads the 36 x 7 cell boundaries, sizes the runs (prefix sum), and has threads 0..71
//      issue one bulk copy per run and array (cp.async.bulk, completion on an mbarrier):
//      ~1700 records x 64 B arrive in shared memory without passing through registers or L1;
//   2. derives the FP32 coordinates of the staged records relative to the tile (the
//      pre-filter of the sweep) - once per tile instead of once per (particle, candidate);
//   3. two warps per own cell, groups of <= 8 own particles: PHASE A - each warp takes every
//      second of the 25 cell columns around the cell and tests their candidates with lanes =
//      (own particle) x (candidate slot), so one shared-memory read of a candidate serves
//      every own particle; survivors go to per-lane hit lists. PHASE B - each warp takes every
//      second own particle and walks BOTH warps' lists of it, 32 pairs per trip with full
//      lanes, records read from shared memory, the particle's own state in registers, exact
//      FP64 membership test |r_a - r_b|^2 <= (2h)^2 as everywhere (geom/bsphere.hpp:52-53);
//   4. one lane per own particle finishes it. No atomics; the order of every sum is fixed by
//      the sorted arrays, not by the launch configuration.
// Against the gather traversal (warp_neighbors: every warp sweeps ~700 candidates for ONE
// particle through L1) this removes the 8-fold repeated sweep of a cell's particles, the L1
// gather traffic of the hit records and most of the per-particle set-up.
// Tiles whose 216 cells hold more records than fit (kTileCap) take the gather traversal.
#pragma once

namespace titgpu {

constexpr int kTileWarps = 16;     // two warps per own cell
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kTileCap = 2048;     // staged records per tile (lattice: 216 cells x 8 = 1728)
constexpr int kTileListCap = 48;   // hit-list entries per lane (lattice: 256 hits over 2 warps x >= 4 slots)
constexpr int kTileCols = 36;      // 6 x 6 cell columns of the staged region
constexpr int kTileGroup = 8;      // own particles handled together (lanes = particle x slot): >= 4 slots

struct TileSmem {
  double4 A[kTileCap];
  double4 B[kTileCap];
  float Fx[kTileCap + 8], Fy[kTileCap + 8], Fz[kTileCap + 8];  // FP32 coordinates of the staged records (cell units, relative to the region)
  unsigned short list[kTileWarps][kTileListCap * 32];
  int pre[kTileCols][8];          // prefetched global cell boundaries of the NEXT tile's region
  int lstart[kTileCols][8];       // local slot of the first record of region cell (column, z), z = 0..6
  int gstart[kTileCols];          // global index of the first staged record of the column
  int roff[kTileCols];            // local slot of the first staged record of the column
  int wtot[kTileWarps][kTileGroup];  // hits of own particle p in this warp's lists
  int ovf[kTileWarps];               // a list of this warp overflowed (the group takes the gather traversal)
  int soff[kTileWarps][32];       // per lane: exclusive offset of its list among the lists of the same own particle in this warp
  unsigned long long bar;
  int tile, total;
};
static_assert(sizeof(TileSmem) <= 227 * 1024, "tile staging exceeds the shared memory of an SM");

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Generic-proxy accesses of shared memory (ld / st) before, async-proxy writes (bulk copies) after.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Tiles that hold fluid particles, in tile order within a warp's 32 tiles (the list order only steers which
// SM takes which tile; every particle's sums are formed inside its own tile).
static __global__ void k_tile_list(const unsigned char* __restrict__ cell_fluid, GridDesc g, int ntx, int nty, int ntz, int* __restrict__ list, int* __restrict__ count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int ntiles = ntx * nty * ntz;
  bool take = false;
  if (t < ntiles) {
    const int tz = t % ntz, ty = (t / ntz) % nty, tx = t / (ntz * nty);
    // tiles without a fluid particle have nothing to sum (cell_fluid is set by k_reorder)
    const int z0 = 2 * tz, z1 = min(2 * tz + 2, g.nc[2]);
    for (int dx = 0; dx < 2 && !take; ++dx)
      for (int dy = 0; dy < 2 && !take; ++dy) {
        const int cx = 2 * tx + dx, cy = 2 * ty + dy;
        if (cx < g.nc[0] && cy < g.nc[1]) {
          const int base = col_base<3>(g, cx, cy);
          for (int z = z0; z < z1; ++z) take = take || cell_fluid[base + z] != 0;
        }
      }
  }
  const unsigned m = __ballot_sync(kFull, take);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(kFull, base, 0);
    if (take) list[base + __popc(m & ((1u << lane) - 1u))] = t;
  }
}

// The 36 x 7 cell boundaries of a tile's region, fetched ahead of time into T.pre by
// asynchronous 4-byte copies (LDGSTS) so that their latency hides behind the previous tile.
__device__ __forceinline__ void tile_prefetch(const Dev<3>& S, TileSmem& T, int tflat, int nty, int ntz) {
  const GridDesc& g = S.P.grid;
  const int tid = threadIdx.x;
  if (tid < kTileCols * 7) {
    const int tz = tflat % ntz, ty = (tflat / ntz) % nty, tx = tflat / (ntz * nty);
    const int rc = tid / 7, zz = tid % 7;
    const int cx = 2 * tx - 2 + rc / 6, cy = 2 * ty - 2 + rc % 6;
    if (cx >= 0 && cx < g.nc[0] && cy >= 0 && cy < g.nc[1]) {
      const int cz = min(max(2 * tz - 2 + zz, 0), g.nc[2]);  // == nc[2]: the end of the column
      const int* src = S.cell_start + (col_base<3>(g, cx, cy) + cz);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&T.pre[rc][zz])), "l"(src) : "memory");
    } else {
      T.pre[rc][zz] = 0;  // a column outside the grid: empty
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Steps 1 and 2 of a tile: region table (from the prefetched boundaries), bulk copies, FP32
// coordinates. Returns the number of staged records, or -1 if they do not fit (nothing
// staged). All threads of the block call it; on return the staged data are visible to all.
__device__ __forceinline__ int tile_stage(const Dev<3>& S, TileSmem& T, int tflat, int nty, int ntz, unsigned& phase, bool want_B) {
  const GridDesc& g = S.P.grid;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tz = tflat % ntz, ty = (tflat / ntz) % nty, tx = tflat / (ntz * nty);
  const int c0x = 2 * tx - 2, c0y = 2 * ty - 2, c0z = 2 * tz - 2;  // first cell of the 6 x 6 x 6 region
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();  // T.pre complete and visible; the previous tile's shared memory is free
  if (tid < 32) {
    // lane c (and c + 32 for the last four columns): run length, exclusive prefix, local cell starts
    const int l0 = T.pre[lane][6] - T.pre[lane][0];
    const int l1 = lane < kTileCols - 32 ? T.pre[lane + 32][6] - T.pre[lane + 32][0] : 0;
    int i0 = l0, i1 = l1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t0 = __shfl_up_sync(kFull, i0, o), t1 = __shfl_up_sync(kFull, i1, o);
      if (lane >= o) { i0 += t0; i1 += t1; }
    }
    const int tot0 = __shfl_sync(kFull, i0, 31), tot1 = __shfl_sync(kFull, i1, 31);
    const int total = tot0 + tot1;
    auto put = [&](int rc, int off) {
      const int g0 = T.pre[rc][0];
      T.gstart[rc] = g0;
      T.roff[rc] = off;
#pragma unroll
      for (int zz = 0; zz < 7; ++zz) T.lstart[rc][zz] = off + (T.pre[rc][zz] - g0);
    };
    put(lane, i0 - l0);
    if (lane < kTileCols - 32) put(lane + 32, tot0 + i1 - l1);
    if (lane == 0) {
      T.total = total;
      if (total <= kTileCap) mbar_arrive_expect_tx(&T.bar, unsigned(total) * (want_B ? 64u : 32u));
    }
  }
  __syncthreads();
  const int total = T.total;
  if (total > kTileCap) return -1;
  if (tid < 2 * kTileCols) {
    const int rc = tid % kTileCols, which = tid / kTileCols;
    const int len = (rc + 1 < kTileCols ? T.roff[rc + 1] : total) - T.roff[rc];
    if (len > 0 && (which == 0 || want_B)) {
      fence_proxy_async();
      if (which == 0) bulk_g2s(T.A + T.roff[rc], S.A + T.gstart[rc], unsigned(len) * 32u, &T.bar);
      else bulk_g2s(T.B + T.roff[rc], S.B + T.gstart[rc], unsigned(len) * 32u, &T.bar);
    }
  }
  while (!mbar_try_wait(&T.bar, phase)) {}
  phase ^= 1u;
  // FP32 coordinates in cell units relative to the region's first cell.
  const double ox = g.org[0] + double(c0x) / g.cinv, oy = g.org[1] + double(c0y) / g.cinv, oz = g.org[2] + double(c0z) / g.cinv;
  for (int j = tid; j < total; j += kTileThreads) {
    const double4 a = T.A[j];
    T.Fx[j] = float((a.x - ox) * g.cinv);
    T.Fy[j] = float((a.y - oy) * g.cinv);
    T.Fz[j] = float((a.z - oz) * g.cinv);
  }
  __syncthreads();
  return total;
}

// Lane layout of a group of `np` (<= 16) own particles: Pn = 2^logP >= np particles x Sn = 32 / Pn slots.
struct TileLanes {
  int logP, Pn, Sn, p, sl;
  __device__ __forceinline__ TileLanes(int np, int lane) {
    logP = np <= 1 ? 0 : 32 - __clz(np - 1);
    Pn = 1 << logP;
    Sn = 32 >> logP;
    p = lane & (Pn - 1);
    sl = lane >> logP;
  }
};

// PHASE A for one group: the candidates of this warp's columns (every second of the 25, from
// `half`) against the lane's own particle (FP32 distance in cell units relative to the region);
// hits are appended to the lane's list. Slot s takes every Sn-th ALIGNED BLOCK OF 4 staged
// records of a column: one 16-byte read per coordinate serves 4 candidates and the 8 lanes of
// a slot alike, the hits of a list come in runs of consecutive records (conflict-free reads in
// phase B, see there), and the slots' lists stay balanced. Returns the lane's hit count, or
// -1 if its list overflowed (the group then takes the gather traversal).
__device__ __forceinline__ int tile_sweep(const TileSmem& T, unsigned list_s, const TileLanes& L, int ax, int ay, int az, int half, float px, float py, float pz, int own, bool pact, float thr) {
  const unsigned lane = threadIdx.x & 31;
  unsigned lp = list_s + 2u * lane;                               // shared-space address of the next entry
  const unsigned lend = lp + 64u * unsigned(kTileListCap - 4);    // room for 4 more entries below this
  bool ovf = false;
  for (int col = half; col < 25; col += 2) {
    const int rc = (ax + col / 5) * 6 + (ay + col % 5);
    const int s0 = T.lstart[rc][az], s1 = T.lstart[rc][az + 5];
    for (int jb = (s0 & ~3) + 4 * L.sl; jb < s1; jb += 4 * L.Sn) {
      if (lp > lend) { ovf = true; break; }
      const float4 x4 = *reinterpret_cast<const float4*>(T.Fx + jb), y4 = *reinterpret_cast<const float4*>(T.Fy + jb), z4 = *reinterpret_cast<const float4*>(T.Fz + jb);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
      const bool inner = pact && jb >= s0 && jb + 4 <= s1;  // (the first and the last block of a column may be partial)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dx = px - xs[k], dy = py - ys[k], dz = pz - zs[k];
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const int j = jb + k;
        // (the particle itself adds nothing to the pair sums: left out)
        const bool hit = !(d2 > thr) && j != own && (inner || (pact && j >= s0 && j < s1));
        if (hit) {
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(lp), "h"((unsigned short)j) : "memory");
          lp += 64u;
        }
      }
    }
  }
  return ovf ? -1 : int((lp - (list_s + 2u * lane)) >> 6);
}

// Ghosts and wall particles of the tile pass: their records go to the output buffers unchanged.
template<int D>
__global__ void k_rhs_passthrough(Dev<D> S, RhsArgs A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= S.P.n) return;
  const int oa = S.orig[a];
  if (oa >= S.P.n_owned) rhs_passthrough<D>(S, A, a, oa, Pack<D>::state(S.A, S.B, a));
}

template<int KID, int EOSK>
__global__ void __launch_bounds__(kTileThreads, 1) k_rhs_tile(Dev<3> S, RhsArgs A, const int* __restrict__ tile_list, const int* __restrict__ n_tiles_ptr, int* __restrict__ cursor, int nty, int ntz) {
  constexpr int D = 3;
  static_assert(EOSK != 0, "the tile pass recomputes the neighbours' EOS values from rho (Tait xi = 7 or linear)");
  extern __shared__ __align__(128) unsigned char tile_smem_raw[];
  TileSmem& T = *reinterpret_cast<TileSmem*>(tile_smem_raw);
  const Params& P = S.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) mbar_init(&T.bar, 1);
  __syncthreads();
  unsigned phase = 0;
  const int n_tiles = *n_tiles_ptr;
  double f2max = 0.0;
  unsigned short* list = T.list[warp];
  // Tiles are dealt round-robin (a block's tiles are spread over the whole list, so the fluid-full
  // and the sparse ones average out); the next tile's cell boundaries are fetched a tile ahead.
  int ti = blockIdx.x;
  if (ti < n_tiles) tile_prefetch(S, T, tile_list[ti], nty, ntz);
  for (; ti < n_tiles; ti += gridDim.x) {
    const int tflat = tile_list[ti];
    const int total = tile_stage(S, T, tflat, nty, ntz, phase, true);
    if (ti + int(gridDim.x) < n_tiles) tile_prefetch(S, T, tile_list[ti + gridDim.x], nty, ntz);
    const int cell = warp >> 1, half = warp & 1;
    const int ax = cell >> 2, ay = (cell >> 1) & 1, az = cell & 1;  // own cell inside the tile
    const int rc_own = (2 + ax) * 6 + (2 + ay);
    if (total < 0) {
      // Too many records for the staging buffers: the gather traversal, one warp per particle.
      HitList& H = reinterpret_cast<HitList*>(tile_smem_raw)[warp];
      const int tz = tflat % ntz, ty = (tflat / ntz) % nty, tx = tflat / (ntz * nty);
      const int cx = 2 * tx + ax, cy = 2 * ty + ay, cz = 2 * tz + az;
      if (cx < P.grid.nc[0] && cy < P.grid.nc[1] && cz < P.grid.nc[2]) {
        const int flat = (cx * P.grid.nc[1] + cy) * P.grid.nc[2] + cz;
        for (int a = S.cell_start[flat] + half; a < S.cell_start[flat + 1]; a += 2) {
          const int oa = S.orig[a];
          if (oa >= P.n_owned) continue;
          f2max = fmax(f2max, rhs_particle<D, KID, EOSK>(S, A, H, a, oa, Pack<D>::state(S.A, S.B, a)));
        }
      }
      continue;
    }
    const int own0 = T.lstart[rc_own][2 + az], own1 = T.lstart[rc_own][3 + az];
    const int gbase = T.gstart[rc_own] - T.roff[rc_own];  // global index = local slot + gbase
    const int w0 = warp & ~1;                              // the cell's first warp
    for (int g0 = own0; g0 < own1; g0 += kTileGroup) {
      const int np = min(kTileGroup, own1 - g0);
      const TileLanes L(np, lane);
      const bool pact = L.p < np;
      const int own = g0 + (pact ? L.p : 0);
      const int oa = pact ? S.orig[own + gbase] : 0x7fffffff;
      const bool fluid = pact && oa < P.n_owned;  // (ghosts and wall particles pass through in k_rhs_passthrough)
      const unsigned fmask = __ballot_sync(kFull, fluid && L.sl == 0);  // bit q: own particle q is an owned fluid particle
      if (fmask == 0) continue;  // (the same decision in both warps of the cell)
      // PHASE A
      const int cnt = tile_sweep(T, smem_u32(list), L, ax, ay, az, half, T.Fx[own], T.Fy[own], T.Fz[own], own, fluid, P.pre_thr);
      const bool ovf = __any_sync(kFull, cnt < 0);
      // Exclusive offset of this lane's list among the lists of its own particle (slots ascending),
      // and the particle's total, per warp.
      int incl = max(cnt, 0);
      for (int o = L.Pn; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      T.soff[warp][lane] = incl - max(cnt, 0);
      if (L.sl == L.Sn - 1 && pact) T.wtot[warp][L.p] = incl;
      if (lane == 0) T.ovf[warp] = ovf;
      named_barrier(1 + cell, 64);  // both warps' lists are complete
      if (T.ovf[w0] | T.ovf[w0 + 1]) {
        // A list overflowed (a very dense neighbourhood): this group by the gather traversal.
        HitList& H = *reinterpret_cast<HitList*>(list);
        static_assert(sizeof(HitList) <= sizeof(T.list[0]), "the gather traversal's scratch must fit a warp's hit lists");
        named_barrier(1 + cell, 64);  // nobody reads the lists any more
        for (int q = half; q < np; q += 2) {
          if (!((fmask >> q) & 1u)) continue;
          const int a = g0 + q + gbase;
          f2max = fmax(f2max, rhs_particle<D, KID, EOSK>(S, A, H, a, S.orig[a], Pack<D>::state(S.A, S.B, a)));
        }
        __syncwarp();
        continue;
      }
      // PHASE B: this warp's own particles (every second one), over both warps' lists.
      double tot_c = 0.0;
      Vec<D> tot_m = vzero<D>();
      for (int q = half; q < np; q += 2) {
        if (!((fmask >> q) & 1u)) continue;
        const int n0 = T.wtot[w0][q], nq = n0 + T.wtot[w0 + 1][q];
        const double4 ar = T.A[g0 + q], br = T.B[g0 + q];
        Vec<D> ra, va;
        ra[0] = ar.x; ra[1] = ar.y; ra[2] = ar.z;
        va[0] = br.x; va[1] = br.y; va[2] = br.z;
        const double rho_a = ar.w;
        double cs_a, Pa, irho_a;
        eos_of_neighbor<EOSK>(P, rho_a, cs_a, Pa, irho_a);
        const double K_a = 2.0 * P.mu / rho_a;
        double pc = 0.0;
        Vec<D> pm = vzero<D>();
        const int hf = (lane >> 2) & 1;
#pragma unroll 1
        for (int b0 = 0; b0 < nq; b0 += 32) {
          const bool act = b0 + lane < nq;
          int h = min(b0 + lane, nq - 1);  // (idle lanes repeat the last hit with weight 0: no branches)
          const int w = h < n0 ? w0 : w0 + 1;
          h -= h < n0 ? 0 : n0;
          // which of the particle's Sn lists of warp w holds hit h: binary search over the slots
          int sidx = 0;
          for (int step = L.Sn >> 1; step > 0; step >>= 1) sidx += h >= T.soff[w][q + ((sidx + step) << L.logP)] ? step : 0;
          const int src = q + (sidx << L.logP);
          const int j = T.list[w][(h - T.soff[w][src]) * 32 + src];
          // Records are 32 bytes, shared-memory loads 16: lanes 0-3 of every 8 read the first
          // half first, lanes 4-7 the second - runs of 4 consecutive records then touch all 32 banks once.
          const double2* pa_ = reinterpret_cast<const double2*>(T.A + j);
          const double2* pb_ = reinterpret_cast<const double2*>(T.B + j);
          const double2 a0 = pa_[hf], a1 = pa_[hf ^ 1], b0_ = pb_[hf], b1_ = pb_[hf ^ 1];
          PState<D> sb;
          sb.r[0] = hf ? a1.x : a0.x; sb.r[1] = hf ? a1.y : a0.y; sb.r[2] = hf ? a0.x : a1.x; sb.rho = hf ? a0.y : a1.y;
          sb.v[0] = hf ? b1_.x : b0_.x; sb.v[1] = hf ? b1_.y : b0_.y; sb.v[2] = hf ? b0_.x : b1_.x; sb.m = hf ? b0_.y : b1_.y;
          rhs_pair<D, KID, EOSK>(P, ra, va, rho_a, cs_a, Pa, K_a, sb, make_double4(0.0, 0.0, 0.0, 0.0), act, pc, pm);
        }
        pc = warp_sum(pc);
        pm = warp_sum(pm);
        if (lane == q) { tot_c = pc; tot_m = pm; }
      }
      // one lane per own particle finishes it
      if (lane < np && (lane & 1) == half && ((fmask >> lane) & 1u)) {
        const int a = g0 + lane + gbase;
        const double4 ar = T.A[g0 + lane], br = T.B[g0 + lane];
        Vec<D> ra, va;
        ra[0] = ar.x; ra[1] = ar.y; ra[2] = ar.z;
        va[0] = br.x; va[1] = br.y; va[2] = br.z;
        f2max = fmax(f2max, rhs_finish<D>(S, A, a, oa, ra, va, ar.w, br.w, S.C[a].x, tot_c, tot_m));
      }
      named_barrier(1 + cell, 64);  // the lists may be rewritten by the next group
    }
  }
  if (A.track_fmax) {
    f2max = warp_max(f2max);
    if (lane == 0 && f2max > 0.0) atomicMax(A.fmax_bits, (unsigned long long)__double_as_longlong(f2max));
  }
}

}  // namespace titgpu
