// Host-side check of the GEMM tile rasterisation (csrc/gemm_common.cuh: tile_mn): for every geometry the persistent
// schedule must visit each (m, n) output tile exactly once, in every mode (plain, L2-grouped, rank-shifted, and the
// "local rows first" order of the in-kernel all-gather).  Built and run by tests/test_tile_order_cpu.py with nvcc.
#include <cstdio>
#include <vector>

#include "gemm_common.cuh"

using namespace dtg;

static int check(int num_m, int num_n, GemmDist d, int local_m_tiles, const char* what) {
  std::vector<int> seen(num_m * num_n, 0);
  for (int t = 0; t < num_m * num_n; ++t) {
    int m = -1, n = -1;
    tile_mn(t, num_m, d, local_m_tiles, m, n);
    if (m < 0 || m >= num_m || n < 0 || n >= num_n) {
      std::printf("FAIL %s: tile %d -> (%d, %d) outside %d x %d\n", what, t, m, n, num_m, num_n);
      return 1;
    }
    seen[m * num_n + n]++;
  }
  for (int i = 0; i < num_m * num_n; ++i)
    if (seen[i] != 1) {
      std::printf("FAIL %s: tile (%d, %d) visited %d times (%d x %d, group_m %d)\n", what, i / num_n, i % num_n, seen[i],
                  num_m, num_n, d.group_m);
      return 1;
    }
  return 0;
}

int main() {
  int bad = 0, cases = 0;
  for (int num_m : {1, 2, 7, 16, 48, 86, 125})
    for (int num_n : {1, 3, 16, 43, 125}) {
      for (int group_m : {0, 1, 4, 8, 20, 32, 200}) {  // plain GEMM: L2-aware groups
        GemmDist d{};
        d.group_m = group_m;
        d.num_n_tiles = num_n;
        bad += check(num_m, num_n, d, 0, "grouped");
        ++cases;
      }
      for (int shift = 0; shift < num_m; shift += (num_m > 4 ? num_m / 4 : 1)) {  // tensor-parallel: start on own rows
        GemmDist d{};
        d.m_tile_shift = shift;
        d.num_n_tiles = num_n;
        bad += check(num_m, num_n, d, 0, "shifted");
        ++cases;
      }
      for (int nranks : {2, 4, 8}) {  // in-kernel all-gather: local row tiles first, then the fetched ones
        if (num_m % nranks) continue;
        const int local = num_m / nranks;
        for (int rank = 0; rank < nranks; ++rank) {
          GemmDist d{};
          d.m_tile_shift = rank * local;
          d.k_shift = num_n;  // tile_mn reads the N tile count from this field in that mode
          d.num_n_tiles = num_n;
          bad += check(num_m, num_n, d, local, "all-gather order");
          ++cases;
        }
      }
    }
  std::printf("%s: %d geometries\n", bad ? "FAILED" : "ok", cases);
  return bad ? 1 : 0;
}
