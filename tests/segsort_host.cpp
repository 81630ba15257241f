// Host harness for csrc/segsort.cuh: the per-row routines of the experimental
// prefix-sorted path, run serially over random inputs and checked against
// std::stable_sort.  Built and run by tests/test_segsort_host.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "segsort.cuh"

int main() {
  std::mt19937_64 rng(12345);
  for (int trial = 0; trial < 200; trial++) {
    const int L = 2 + (int)(rng() % 4);
    const uint64_t n = 1 + rng() % 3000;
    const unsigned max_run = 1 + (unsigned)(rng() % 40);
    std::vector<std::vector<uint64_t>> col(L, std::vector<uint64_t>(n));
    uint64_t f[8] = {0};
    for (int l = 0; l < L; l++) f[l] = (rng() & 1) ? 0x8000000000000000ull : 0ull;
    // lane 0: non-decreasing in flipped order, runs of random length <= max_run
    uint64_t v = f[0] ? (uint64_t)(-5000) : 0;   // signed lanes start negative
    for (uint64_t i = 0; i < n;) {
      uint64_t len = 1 + rng() % max_run;
      for (uint64_t k = 0; k < len && i < n; k++, i++) col[0][i] = v;
      v += 1 + rng() % 3;
    }
    for (int l = 1; l < L; l++)
      for (uint64_t i = 0; i < n; i++) col[l][i] = (rng() % 7) - (f[l] ? 3 : 0);   // few values: ties; negative when signed
    const uint64_t* c[8];
    for (int l = 0; l < L; l++) c[l] = col[l].data();
    // census
    uint64_t inv = 0;
    unsigned longest = 0;
    for (uint64_t i = 0; i < n; i++) {
      unsigned run;
      inv += seg_lane0_props(c[0], f[0], n, i, &run);
      longest = std::max(longest, run);
    }
    if (inv != 0) { printf("trial %d: lane 0 reported unsorted\n", trial); return 1; }
    unsigned true_longest = 0;
    for (uint64_t i = 0; i < n;) { uint64_t j = i; while (j < n && col[0][j] == col[0][i]) j++; true_longest = std::max<unsigned>(true_longest, (unsigned)(j - i)); i = j; }
    if (std::min(true_longest, SEG_RUN_CAP + 1) != longest) { printf("trial %d: run %u vs %u\n", trial, longest, true_longest); return 1; }
    // rank
    std::vector<uint32_t> idx(n, 0xffffffffu);
    for (uint64_t i = 0; i < n; i++) seg_rank_row(c, f, L, n, i, idx.data());
    std::vector<uint32_t> want(n);
    for (uint64_t i = 0; i < n; i++) want[i] = (uint32_t)i;
    std::stable_sort(want.begin(), want.end(), [&](uint32_t a, uint32_t b) {
      for (int l = 0; l < L; l++) {
        uint64_t x = col[l][a] ^ f[l], y = col[l][b] ^ f[l];
        if (x != y) return x < y;
      }
      return false;
    });
    if (idx != want) { printf("trial %d: permutation differs (L=%d n=%llu)\n", trial, L, (unsigned long long)n); return 1; }
    // an inversion on lane 0 must be seen
    if (n >= 2) {
      std::swap(col[0][0], col[0][n - 1]);
      uint64_t inv2 = 0;
      for (uint64_t i = 0; i < n; i++) { unsigned run; inv2 += seg_lane0_props(c[0], f[0], n, i, &run); }
      if ((col[0][0] != col[0][n - 1]) && inv2 == 0) { printf("trial %d: inversion missed\n", trial); return 1; }
    }
  }
  printf("segsort host harness: OK\n");
  return 0;
}
