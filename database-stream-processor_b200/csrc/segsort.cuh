// segsort.cuh — sort of rows whose first lane is already non-decreasing
// (EXPERIMENTAL, off unless DBSP_PREFIX_SORT is set; see consolidate.cu).
//
// Time-ordered event tables (Nexmark bids keyed by date_time) reach
// Batch::from_tuples (trace/mod.rs:259-263 -> consolidation/mod.rs:32-52) with
// lane 0 in order and only the minor lanes of each equal-lane-0 run (a handful
// of rows) unordered.  Instead of a full multi-word LSD radix sort, every row
// computes its rank inside its run — O(run^2) compares over rows that sit in
// L1 — and the sorted order is written as a row-id permutation, the same form
// the radix path hands to the duplicate/zero epilogue.
//
// The per-row routines are plain functions so that the logic is unit-tested
// on the host (tests/test_segsort_host.py compiles them with g++).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SEG_HD __host__ __device__ __forceinline__
#else
#define SEG_HD inline
#endif

constexpr unsigned SEG_RUN_CAP = 32;   // longest run the rank kernel is used for

// Row i of lane 0 (values already order-flipped with f0): returns 1 when lane 0
// decreases from row i-1 to row i.  *run_len = length of the equal-lane-0 run
// that starts at i (counted up to SEG_RUN_CAP + 1), 0 when i is not a run head.
SEG_HD unsigned seg_lane0_props(const uint64_t* lane0, uint64_t f0, uint64_t n, uint64_t i, unsigned* run_len) {
  const uint64_t a = lane0[i] ^ f0;
  unsigned inv = 0;
  bool head = true;
  if (i > 0) {
    const uint64_t p = lane0[i - 1] ^ f0;
    inv = p > a ? 1u : 0u;
    head = p != a;
  }
  unsigned len = 0;
  if (head) {
    len = 1;
    while (len <= SEG_RUN_CAP && i + len < n && (lane0[i + len] ^ f0) == a) len++;
  }
  *run_len = len;
  return inv;
}

// Row i finds its run [s, e) of equal lane 0, counts the rows of the run that
// sort before it on lanes 1..L-1 (ties broken by position: stable) and records
// itself at its sorted position: idx[s + rank] = i.
SEG_HD void seg_rank_row(const uint64_t* const* c, const uint64_t* f, int L, uint64_t n, uint64_t i, uint32_t* idx) {
  const uint64_t* lane0 = c[0];
  const uint64_t a = lane0[i];
  uint64_t s = i, e = i + 1;
  while (s > 0 && lane0[s - 1] == a) s--;
  while (e < n && lane0[e] == a) e++;
  uint64_t mine[8];
  for (int l = 1; l < L; l++) mine[l] = c[l][i] ^ f[l];
  uint32_t rank = 0;
  for (uint64_t j = s; j < e; j++) {
    if (j == i) continue;
    int cmp = 0;   // row j vs row i on the minor lanes
    for (int l = 1; l < L; l++) {
      const uint64_t x = c[l][j] ^ f[l];
      if (x != mine[l]) { cmp = x < mine[l] ? -1 : 1; break; }
    }
    if (cmp < 0 || (cmp == 0 && j < i)) rank++;
  }
  idx[s + rank] = (uint32_t)i;
}
