// Host-side check of the cell order helpers of csrc/engine.cuh (col_base / cell_flat / cell_unflat are
// __host__ __device__): compiled and run by tests/test_cell_order.py, no GPU needed.
#include <cstdio>
#include <vector>

#include "engine.cuh"

using namespace titgpu;

static int check(int nx, int ny, int nz, int tyl) {
  GridDesc g{};
  g.nc[0] = nx; g.nc[1] = ny; g.nc[2] = nz;
  g.tyl = tyl;
  const int T = 1 << tyl;
  const long long nyp = tyl ? ((ny + T - 1) / T) * (long long)T : ny;
  g.ncells = int(nx * nyp * nz);
  std::vector<char> seen(size_t(g.ncells), 0);
  int bad = 0;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      const int base = col_base<3>(g, x, y);
      for (int z = 0; z < nz; ++z) {
        int ci[3] = {x, y, z};
        const int f = cell_flat<3>(g, ci);
        if (f != base + z || f < 0 || f >= g.ncells || seen[size_t(f)]) { ++bad; continue; }  // a column is one contiguous run; the map is injective
        seen[size_t(f)] = 1;
        int cj[3];
        if (!cell_unflat<3>(g, f, cj) || cj[0] != x || cj[1] != y || cj[2] != z) ++bad;
      }
    }
  // the cells that no (x, y, z) maps to are exactly the padding of the last tile
  long long unseen = 0;
  for (int f = 0; f < g.ncells; ++f)
    if (!seen[size_t(f)]) {
      ++unseen;
      int cj[3];
      if (cell_unflat<3>(g, f, cj)) ++bad;
    }
  if (unseen != (long long)nx * (nyp - ny) * nz) ++bad;
  // within a tile, the columns (x, y) .. (x, y + T - 1) of one x are adjacent runs, and x + 1 follows
  if (tyl) {
    for (int x = 0; x + 1 < nx; ++x)
      for (int y = 0; y < ny; ++y) {
        if ((y & (T - 1)) != T - 1 && y + 1 < ny && col_base<3>(g, x, y + 1) - col_base<3>(g, x, y) != nz) ++bad;
        if (col_base<3>(g, x + 1, y) - col_base<3>(g, x, y) != T * nz) ++bad;
      }
  }
  return bad;
}

int main() {
  int bad = 0;
  const int dims[][3] = {{1, 1, 1}, {3, 5, 2}, {7, 16, 3}, {5, 17, 4}, {9, 33, 5}, {4, 64, 1}, {6, 100, 7}};
  for (const auto& d : dims)
    for (int tyl = 0; tyl <= 5; ++tyl) bad += check(d[0], d[1], d[2], tyl);
  // 2-D: plain row-major
  {
    GridDesc g{};
    g.nc[0] = 5; g.nc[1] = 7; g.ncells = 35; g.tyl = 0;
    for (int x = 0; x < 5; ++x)
      for (int y = 0; y < 7; ++y) {
        int ci[2] = {x, y}, cj[2];
        const int f = cell_flat<2>(g, ci);
        if (f != x * 7 + y || !cell_unflat<2>(g, f, cj) || cj[0] != x || cj[1] != y || col_base<2>(g, x, 0) != x * 7) ++bad;
      }
  }
  // thread chunks of the scan kernels: every particle exactly once, whatever the count
  for (int n : {1, 5, 31, 32, 33, 1000, 4096, 4097, 14300, 83748, 131071, 131072, 262143, 262144, 300001}) {
    struct M { int count, per, nchunks, transposed; int at(int chunk, int lane) const { const int t = transposed ? lane * nchunks + chunk : chunk * 32 + lane; return lane < per && t < count ? t : count; } };
    const M m{n, ScanMap::per_chunk(n), ScanMap::chunks(n), n < (1 << 18)};
    std::vector<char> seen(size_t(n), 0);
    for (int c = 0; c < m.nchunks; ++c)
      for (int l = 0; l < 32; ++l) {
        const int t = m.at(c, l);
        if (t == n) continue;
        if (t < 0 || t > n || seen[size_t(t)]) { ++bad; continue; }
        seen[size_t(t)] = 1;
      }
    for (int t = 0; t < n; ++t) bad += !seen[size_t(t)];
  }
  std::printf("%s (%d)\n", bad ? "FAILED" : "ok", bad);
  return bad != 0;
}
