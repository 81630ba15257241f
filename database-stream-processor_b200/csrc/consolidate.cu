// consolidate.cu — K1/K2: Batch::from_tuples on the device.
//
// Replaces consolidation::consolidate (sort + sum equal keys + drop zeros,
// crates/dbsp/src/trace/consolidation/mod.rs:32-52,182-231), the MergeBatcher
// (trace/ord/merge_batcher/mod.rs:65-80,155-197) and the Builder
// (trace/layers/ordered/mod.rs:874-888) for rows of <= 8 integer lanes.
//
// B200-first design.
//  (1) One property pass over the rows reduces each lane's [min,max] and
//      counts order inversions, adjacent duplicates and zero weights.
//  (2) Rows that arrive already ordered (projections of sorted batches, join
//      and gather outputs, id-ordered event tables) skip the sort entirely;
//      if they are also duplicate- and zero-free the input buffer *becomes*
//      the batch (no copy).
//  (3) Otherwise the comparison sort the reference spends "90% of the work"
//      in (consolidation/mod.rs:101-104) is a radix sort over bit-packed
//      composite keys: the lanes' significant bits are concatenated (order
//      preserving, injective) into as few 64-bit words as possible and only
//      those bits are sorted, as (key word, row id) pairs, by the hand-written
//      passes of sort.cu (top digits in HBM, the rest in shared memory; no HBM
//      pass at all when the leading lane arrives ordered).  The Nexmark
//      schemas pack into one word of 30-60 bits.
//  (4) Epilogue: a two-pass reduce-by-key sums runs of equal rows and drops
//      zero sums; the duplicate-free case is one unpack/gather pass.
#include "common.cuh"

namespace {

struct Plan {
  int L, W;
  int use_key;   // epilogue may compare / unpack the single packed word
  u64 mn[MAXL], flip[MAXL], mask[MAXL];
  unsigned char word[MAXL], shift[MAXL], bits[MAXL];
  unsigned char wbits[MAXL];
};

// mm[l] = min, mm[L+l] = max of flipped lane l; mm[2L] = # inversions
// (row i-1 > row i), mm[2L+1] = # adjacent duplicates, mm[2L+2] = # zero weights,
// mm[2L+3] = row count (when it lives on the device), mm[2L+4] = # inversions of lane 0 alone.
__global__ void k_props(Cols cols, Flips f, int L, const i64* w, u64 n_host, const u32* dn, u64* mm) {
  // the producer may have left the exact row count on the device (dn): the
  // census then returns it with the lane ranges in the same read-back
  const u64 n = dn ? (u64)*dn : n_host;
  if (dn && blockIdx.x == 0 && threadIdx.x == 0) mm[2 * L + 3] = n;
  __shared__ u64 smin[MAXL], smax[MAXL];
  __shared__ unsigned s_cnt[4];
  if (threadIdx.x < MAXL) { smin[threadIdx.x] = ~0ull; smax[threadIdx.x] = 0; }
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  u64 lmin[MAXL], lmax[MAXL];
  for (int l = 0; l < L; l++) { lmin[l] = ~0ull; lmax[l] = 0; }
  unsigned inv = 0, dup = 0, zero = 0, inv0 = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    int c = 0;   // cmp(row i-1, row i)
    for (int l = 0; l < L; l++) {
      u64 v = cols.c[l][i] ^ f.f[l];
      lmin[l] = min(lmin[l], v);
      lmax[l] = max(lmax[l], v);
      if (i > 0 && c == 0) {
        u64 pv = cols.c[l][i - 1] ^ f.f[l];
        if (pv != v) c = pv < v ? -1 : 1;
        if (l == 0) inv0 += pv > v;
      }
    }
    if (i > 0) { inv += c > 0; dup += c == 0; }
    if (w && w[i] == 0) zero++;
  }
  for (int l = 0; l < L; l++) {
    u64 a = lmin[l], b = lmax[l];
    for (int o = 16; o > 0; o >>= 1) {
      a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin((unsigned long long*)&smin[l], (unsigned long long)a);
      atomicMax((unsigned long long*)&smax[l], (unsigned long long)b);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    inv += __shfl_xor_sync(0xffffffffu, inv, o);
    dup += __shfl_xor_sync(0xffffffffu, dup, o);
    zero += __shfl_xor_sync(0xffffffffu, zero, o);
    inv0 += __shfl_xor_sync(0xffffffffu, inv0, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (inv) atomicAdd(&s_cnt[0], inv);
    if (dup) atomicAdd(&s_cnt[1], dup);
    if (zero) atomicAdd(&s_cnt[2], zero);
    if (inv0) atomicAdd(&s_cnt[3], inv0);
  }
  __syncthreads();
  if (threadIdx.x < L) {
    atomicMin((unsigned long long*)&mm[threadIdx.x], (unsigned long long)smin[threadIdx.x]);
    atomicMax((unsigned long long*)&mm[L + threadIdx.x], (unsigned long long)smax[threadIdx.x]);
  }
  if (threadIdx.x < 3 && s_cnt[threadIdx.x])
    atomicAdd((unsigned long long*)&mm[2 * L + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
  if (threadIdx.x == 3 && s_cnt[3]) atomicAdd((unsigned long long*)&mm[2 * L + 4], (unsigned long long)s_cnt[3]);
}

__global__ void k_init_props(u64* mm, int L) {
  int t = threadIdx.x;
  if (t < L) mm[t] = ~0ull;
  else if (t < 2 * L + 5) mm[t] = 0;
}

// key[i] = word `wd` of row (idx ? idx[i] : i); writes idx_out[i] = i when idx == nullptr.
__global__ void k_pack(Cols cols, Plan p, int wd, const u32* idx, u64 n, u64* key, u32* idx_out) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 r = idx ? idx[i] : i;
  u64 k = 0;
  for (int l = 0; l < p.L; l++)
    if (p.word[l] == wd && p.bits[l]) k |= ((cols.c[l][r] ^ p.flip[l]) - p.mn[l]) << p.shift[l];
  key[i] = k;
  if (!idx) idx_out[i] = (u32)i;
}

__device__ __forceinline__ bool rows_differ(const Cols& cols, int L, u64 a, u64 b) {
  for (int l = 0; l < L; l++)
    if (cols.c[l][a] != cols.c[l][b]) return true;
  return false;
}

// Head flags (row differs from its predecessor in sorted order) + gathered
// weights; `counters` (optional) accumulate # non-heads and # zero weights.
__global__ void k_heads(Cols cols, int L, int use_key, const u64* key, const u32* idx, const i64* w, u64 n, u32* flags,
                        i64* ws, u64* counters) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned bad = 0;
  if (counters && counters[2]) return;   // the sort raised its fallback flag: key / idx are not a permutation yet
  if (i < n) {
    u64 r = idx ? idx[i] : i;
    bool head = true;
    if (i > 0) head = use_key ? (key[i] != key[i - 1]) : rows_differ(cols, L, r, idx ? idx[i - 1] : i - 1);
    i64 wt = w ? w[r] : 1;
    if (flags) { flags[i] = head ? 1u : 0u; ws[i] = wt; }
    bad = (!head ? 1u : 0u) | (wt == 0 ? 2u : 0u);
  } else if (i == n && flags) {
    flags[n] = 0;
  }
  if (counters) {
    unsigned d = __ballot_sync(0xffffffffu, bad & 1u), z = __ballot_sync(0xffffffffu, bad & 2u);
    if ((threadIdx.x & 31) == 0 && (d | z)) {
      if (d) atomicAdd((unsigned long long*)&counters[0], (unsigned long long)__popc(d));
      if (z) atomicAdd((unsigned long long*)&counters[1], (unsigned long long)__popc(z));
    }
  }
}

__device__ __forceinline__ u64 unpack_lane(const Plan& p, int l, u64 key) {
  u64 v = p.bits[l] ? ((key >> p.shift[l]) & p.mask[l]) : 0;
  return (v + p.mn[l]) ^ p.flip[l];
}

// Duplicate-free fast path: output row i = input row idx[i].
__global__ void k_emit_unique(Cols cols, Plan p, const u64* key, const u32* idx, const i64* w, u64 n, MCols out,
                              i64* out_w) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 r = idx ? idx[i] : i;
  if (p.use_key) {
    u64 k = key[i];
    for (int l = 0; l < p.L; l++) out.c[l][i] = unpack_lane(p, l, k);
  } else {
    for (int l = 0; l < p.L; l++) out.c[l][i] = cols.c[l][r];
  }
  out_w[i] = w ? w[r] : 1;
}

// ---------------------------------------------------------------------------
// Reduce-by-key over rows that are already in sorted order (identity order, or
// through the row ids of the radix sort): every run of equal rows becomes one
// row carrying the sum of the run's weights; zero sums are dropped.  Two data
// passes around a small per-tile scan:
//   pass 0: per tile of RBK_TILE rows — #heads, weight of the rows before the
//           first head (continuation of the run entering the tile), weight
//           since the last head, #kept runs that lie entirely inside the tile;
//   k_rbk_scan: carries the open run's weight across tiles (flag/sum monoid)
//           and turns kept counts into output offsets;
//   pass 1: recomputes the tile with its carry-in and writes the kept runs.
// Replaces the dedup+retain of consolidate (consolidation/mod.rs:32-52).
constexpr int RBK_THREADS = 256, RBK_R = 4, RBK_TILE = RBK_THREADS * RBK_R;
struct RbkTile {
  i64 pre, post;      // weight before the first head / since the last head (all rows if no head)
  u32 nheads, inner_kept, closed_last, pad;
};
struct RbkCarry {
  i64 carry_in;
  u32 out_base, pad;
};
struct HS {
  u32 h;
  i64 s;
};
struct HSOp {
  __device__ __forceinline__ HS operator()(const HS& a, const HS& b) const {
    HS r;
    r.h = a.h | b.h;
    r.s = b.h ? b.s : (i64)((u64)a.s + (u64)b.s);
    return r;
  }
};

__device__ __forceinline__ bool rbk_differs(const Cols& cols, int L, int use_key, const u64* key, const u32* idx, u64 i,
                                            u64 j) {
  if (use_key) return key[i] != key[j];
  u64 a = idx ? idx[i] : i, b = idx ? idx[j] : j;
  for (int l = 0; l < L; l++)
    if (cols.c[l][a] != cols.c[l][b]) return true;
  return false;
}

template <int PHASE>
__global__ void __launch_bounds__(RBK_THREADS)
k_rbk(Cols cols, Plan p, const u64* key, const u32* idx, const i64* w, u64 n, RbkTile* tiles, const RbkCarry* carry,
      MCols out, i64* out_w) {
  __shared__ HS s_warp_hs[RBK_THREADS / 32];
  __shared__ u32 s_warp_u[RBK_THREADS / 32];
  __shared__ i64 s_pre;
  __shared__ u32 s_inner, s_closed;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u64 tile_s = (u64)blockIdx.x * RBK_TILE;
  const u64 tile_e = tile_s + RBK_TILE < n ? tile_s + RBK_TILE : n;
  const u64 r0 = tile_s + (u64)tid * RBK_R;
  if (tid == 0) { s_pre = 0; s_inner = 0; s_closed = 0; }

  // per-row head / tail flags and weights of this thread's rows
  bool hd[RBK_R], tl[RBK_R], valid[RBK_R];
  i64 wt[RBK_R];
#pragma unroll
  for (int k = 0; k < RBK_R; k++) {
    const u64 i = r0 + k;
    valid[k] = i < tile_e;
    hd[k] = false; tl[k] = false; wt[k] = 0;
    if (valid[k]) {
      hd[k] = (i == 0) || rbk_differs(cols, p.L, p.use_key, key, idx, i, i - 1);
      wt[k] = w ? w[idx ? idx[i] : i] : 1;
    }
  }
#pragma unroll
  for (int k = 0; k < RBK_R; k++) {
    const u64 i = r0 + k;
    if (valid[k]) {
      if (k + 1 < RBK_R && r0 + k + 1 < tile_e) tl[k] = hd[k + 1];
      else tl[k] = (i + 1 >= n) || rbk_differs(cols, p.L, p.use_key, key, idx, i + 1, i);
    }
  }
  // thread aggregate: (has head, weight since last head / total)
  HS agg; agg.h = 0; agg.s = 0;
  i64 pre_t = 0;   // weight before this thread's first head
#pragma unroll
  for (int k = 0; k < RBK_R; k++) {
    if (!valid[k]) continue;
    if (hd[k]) { agg.h = 1; agg.s = wt[k]; }
    else { agg.s = (i64)((u64)agg.s + (u64)wt[k]); if (!agg.h) pre_t = agg.s; }
  }
  // block exclusive scan of the aggregates with the flag/sum monoid
  HSOp op;
  HS incl = agg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    HS v;
    v.h = __shfl_up_sync(0xffffffffu, incl.h, o);
    v.s = __shfl_up_sync(0xffffffffu, incl.s, o);
    if (lane >= o) incl = op(v, incl);
  }
  if (lane == 31) s_warp_hs[wid] = incl;
  __syncthreads();
  HS wprefix; wprefix.h = 0; wprefix.s = 0;
  for (int k = 0; k < wid; k++) wprefix = op(wprefix, s_warp_hs[k]);
  HS excl_in_warp;
  excl_in_warp.h = __shfl_up_sync(0xffffffffu, incl.h, 1);
  excl_in_warp.s = __shfl_up_sync(0xffffffffu, incl.s, 1);
  if (lane == 0) { excl_in_warp.h = 0; excl_in_warp.s = 0; }
  const HS excl = op(wprefix, excl_in_warp);   // open run before this thread, within the tile

  const i64 cin = PHASE ? carry[blockIdx.x].carry_in : 0;
  // running weight of the run open at this thread's first row
  i64 run = excl.h ? excl.s : (i64)((u64)excl.s + (u64)cin);
  bool head_seen = excl.h != 0;
  u32 kept = 0, inner = 0;
  i64 tot[RBK_R];
  bool kp[RBK_R];
#pragma unroll
  for (int k = 0; k < RBK_R; k++) {
    kp[k] = false; tot[k] = 0;
    if (!valid[k]) continue;
    if (hd[k]) { run = wt[k]; head_seen = true; }
    else run = (i64)((u64)run + (u64)wt[k]);
    if (tl[k]) {
      tot[k] = run;
      if (PHASE) { kp[k] = run != 0; kept += kp[k]; }
      else if (head_seen) inner += (run != 0);   // a run that lies entirely inside this tile
    }
  }

  if (PHASE == 0) {
    // tile aggregate
    if (!excl.h) {   // no head before this thread in the tile: its leading rows continue the entering run
      i64 c = agg.h ? pre_t : agg.s;
      if (c) atomicAdd((unsigned long long*)&s_pre, (unsigned long long)c);
    }
    if (inner) atomicAdd(&s_inner, inner);
    if (r0 < tile_e && r0 + RBK_R >= tile_e) {   // owner of the tile's last row
      int lastk = (int)(tile_e - 1 - r0);
      bool c = false;
#pragma unroll
      for (int k = 0; k < RBK_R; k++) if (k == lastk) c = tl[k];
      s_closed = c ? 1u : 0u;
    }
    __syncthreads();
    if (tid == RBK_THREADS - 1) {
      HS tile = op(wprefix, incl);   // inclusive over the whole tile (this is the last thread)
      RbkTile t;
      t.pre = s_pre;
      t.post = tile.s;
      t.nheads = tile.h;
      t.inner_kept = s_inner;
      t.closed_last = s_closed;
      t.pad = 0;
      tiles[blockIdx.x] = t;
    }
    return;
  }

  // PHASE 1: ranks of the kept runs and output
  u32 kincl = kept;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 v = __shfl_up_sync(0xffffffffu, kincl, o);
    if (lane >= o) kincl += v;
  }
  if (lane == 31) s_warp_u[wid] = kincl;
  __syncthreads();
  u32 woff = 0;
  for (int k = 0; k < wid; k++) woff += s_warp_u[k];
  u32 pos = carry[blockIdx.x].out_base + woff + kincl - kept;
#pragma unroll
  for (int k = 0; k < RBK_R; k++) {
    if (!kp[k]) continue;
    const u64 i = r0 + k;
    if (p.use_key) {
      const u64 kk = key[i];
      for (int l = 0; l < p.L; l++) out.c[l][pos] = unpack_lane(p, l, kk);
    } else {
      const u64 r = idx ? idx[i] : i;
      for (int l = 0; l < p.L; l++) out.c[l][pos] = cols.c[l][r];
    }
    out_w[pos] = tot[k];
    pos++;
  }
}

// One CTA: carry the open run's weight across tiles and prefix-sum the kept counts.
__global__ void k_rbk_scan(const RbkTile* tiles, u32 ntiles, RbkCarry* carry, u32* total_out) {
  __shared__ HS s_hs[1024];
  __shared__ u32 s_cnt[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const u32 chunk = (ntiles + nt - 1) / nt;
  const u32 lo = min(ntiles, tid * chunk), hi = min(ntiles, lo + chunk);
  HSOp op;
  // element of tile t for the carry recurrence: carry_out = closes ? 0 : (h ? post : carry_in + pre)
  auto elem = [&](const RbkTile& t) {
    HS e;
    const bool ends = t.closed_last != 0;
    e.h = (t.nheads || ends) ? 1u : 0u;
    e.s = ends ? 0 : (t.nheads ? t.post : t.pre);
    return e;
  };
  HS agg; agg.h = 0; agg.s = 0;
  for (u32 t = lo; t < hi; t++) agg = op(agg, elem(tiles[t]));
  s_hs[tid] = agg;
  __syncthreads();
  if (tid == 0) {   // serial exclusive scan over <= 1024 chunk aggregates
    HS run; run.h = 0; run.s = 0;
    for (int k = 0; k < nt; k++) { HS v = s_hs[k]; s_hs[k] = run; run = op(run, v); }
  }
  __syncthreads();
  HS cur = s_hs[tid];
  u32 cnt = 0;
  for (u32 t = lo; t < hi; t++) {
    const RbkTile tl = tiles[t];
    const i64 cin = cur.s;   // weight of the run open at the tile's first row (0 if none)
    carry[t].carry_in = cin;
    const bool ends_here = tl.nheads || tl.closed_last;
    cnt += tl.inner_kept + ((ends_here && (i64)((u64)cin + (u64)tl.pre) != 0) ? 1u : 0u);
    cur = op(cur, elem(tl));
  }
  s_cnt[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    u32 run = 0;
    for (int k = 0; k < nt; k++) { u32 v = s_cnt[k]; s_cnt[k] = run; run += v; }
    *total_out = run;
  }
  __syncthreads();
  u32 base = s_cnt[tid];
  cur = s_hs[tid];
  for (u32 t = lo; t < hi; t++) {
    const RbkTile tl = tiles[t];
    carry[t].out_base = base;
    const i64 cin = cur.s;
    const bool ends_here = tl.nheads || tl.closed_last;
    base += tl.inner_kept + ((ends_here && (i64)((u64)cin + (u64)tl.pre) != 0) ? 1u : 0u);
    cur = op(cur, elem(tl));
  }
}

inline int bits_for(u64 range) { return range == 0 ? 0 : 64 - __builtin_clzll(range); }

}  // namespace

int32_t radix_sort_pairs(Ctx* ctx, u64* ka, u64* kb, u32* ia, u32* ib, u64 n, int bits, int presorted_top_bits, bool force_lsd,
                         unsigned long long* fail, u64** key_out, u32** idx_out, int* hbm_passes);

int32_t consolidate_rows(Ctx* ctx, const dbsp_schema& s, const Cols& cols, const i64* w, u64 n, const BufP* adopt,
                         Batch** out, const u32* d_n) {
  const int L = s.n_key_lanes + s.n_val_lanes;
  if (n == 0) { *out = batch_new_empty(ctx, s); return DBSP_OK; }   // with d_n: n is an upper bound
  if (n >= (1ull << 32)) { set_error("consolidate: more than 2^32-1 rows in one batch"); return DBSP_ERR_UNSUPPORTED; }
  cudaStream_t st = ctx->stream;
  const int TB = 256;

  Flips f;
  for (int l = 0; l < MAXL; l++) f.f[l] = (l < L && s.lane_types[l] == DBSP_I64) ? 0x8000000000000000ull : 0;

  // ---- (1) lane ranges + order / duplicate / zero-weight census ---------------
  u64 mm[2 * MAXL + 5];
  {
    u64* dmm = ctx->d_scratch + 64;
    k_init_props<<<1, 32, 0, st>>>(dmm, L);
    int g = (int)std::min<u64>((n + TB - 1) / TB, (u64)ctx->sm_count * 8);
    {
      ProfScope ps(ctx, KID_MINMAX, n * (u64)(L + (w ? 1 : 0)) * 8);
      k_props<<<g, TB, 0, st>>>(cols, f, L, w, n, d_n, dmm);
    }
    ctx->kernel_launches += 2;
    TRY(read_back(ctx, dmm, 2 * L + 5, mm));
    if (d_n) {
      n = mm[2 * L + 3];
      if (n == 0) { *out = batch_new_empty(ctx, s); return DBSP_OK; }
    }
  }
  const unsigned nblk = (unsigned)((n + TB - 1) / TB);
  const u64 n_inv = mm[2 * L], n_dup = mm[2 * L + 1], n_zero = mm[2 * L + 2];

  Plan p;
  memset(&p, 0, sizeof(p));
  p.L = L;
  u32* idx_cur = nullptr;
  const u64* key_sorted = nullptr;
  BufP kbuf, ibuf, tmp;

  if (n_inv == 0) {
    // ---- (2) already ordered: no sort ------------------------------------------
    if (n_dup == 0 && n_zero == 0 && adopt && w) {
      Batch* b = new Batch();   // the input buffer *is* the batch
      b->s = s;
      b->n = n;
      b->ctx = ctx;
      for (int l = 0; l < L; l++) b->col[l] = cols.c[l];
      b->w = w;
      b->bufs.push_back(*adopt);
      *out = b;
      return DBSP_OK;
    }
    p.W = 1;
    p.use_key = 0;   // compare / copy the lanes themselves, identity order
  } else {
    // ---- (3) bit-packing plan: lanes from last (least significant) to first ----
    int word = 0, used = 0;
    for (int l = L - 1; l >= 0; l--) {
      int b = bits_for(mm[L + l] - mm[l]);
      if (used + b > 64) { word++; used = 0; }
      p.word[l] = (unsigned char)word;
      p.shift[l] = (unsigned char)used;
      p.bits[l] = (unsigned char)b;
      p.mn[l] = mm[l];
      p.flip[l] = f.f[l];
      p.mask[l] = b >= 64 ? ~0ull : ((1ull << b) - 1);
      used += b;
      p.wbits[word] = (unsigned char)used;
    }
    p.W = word + 1;
    p.use_key = p.W == 1;
    TRY(dev_alloc(ctx, (size_t)n * 8 * 2, &kbuf));
    TRY(dev_alloc(ctx, (size_t)n * 4 * 2, &ibuf));
  }
  u64* const kbufs[2] = {kbuf ? (u64*)kbuf->p : nullptr, kbuf ? (u64*)kbuf->p + n : nullptr};
  u32* const ibufs[2] = {ibuf ? (u32*)ibuf->p : nullptr, ibuf ? (u32*)ibuf->p + n : nullptr};
  u64* const cnt = ctx->d_scratch + 32;   // [0] duplicates, [1] zero weights, [2] sort-fallback flag
  // lane 0 ordered on arrival (event tables come in time order) and the key is one word: its top bits[0] bits are
  // non-decreasing already, the sort can skip every HBM pass (sort.cu)
  const int presorted = (p.W == 1 && n_inv != 0 && mm[2 * L + 4] == 0) ? (int)p.bits[0] : 0;

  // sort (key word, row id) pairs word by word, least significant word first (every stage is stable)
  auto sort_rows = [&](bool force_lsd) -> int32_t {
    idx_cur = nullptr;
    for (int wd = 0; wd < p.W; wd++) {
      // the keys of the previous word are dead: always pack into key buffer 0; the row ids stay where the
      // previous word's sort left them
      u64* kdst = kbufs[0];
      u32* iother;
      {
        ProfScope ps(ctx, KID_PACK, n * (u64)L * 8 + n * 12);
        if (wd == 0) {
          k_pack<<<nblk, TB, 0, st>>>(cols, p, wd, nullptr, n, kdst, ibufs[0]);
          idx_cur = ibufs[0];
        } else {
          k_pack<<<nblk, TB, 0, st>>>(cols, p, wd, idx_cur, n, kdst, nullptr);
        }
      }
      LAUNCH_COUNT(ctx);
      iother = idx_cur == ibufs[0] ? ibufs[1] : ibufs[0];
      key_sorted = kdst;
      if (p.wbits[wd] > 0 && n > 1) {
        u64* ko;
        u32* io;
        TRY(radix_sort_pairs(ctx, kdst, kbufs[1], idx_cur, iother, n, (int)p.wbits[wd], presorted, force_lsd,
                             (unsigned long long*)(cnt + 2), &ko, &io, nullptr));
        key_sorted = ko;
        idx_cur = io;
      }
    }
    return DBSP_OK;
  };
  if (n_inv != 0) {
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 24, st));
    TRY(sort_rows(false));
  }

  // ---- (4) epilogue -------------------------------------------------------------
  // After a sort, duplicates are possible iff the input had any equal pair at
  // all — unknown from the census (it only saw adjacent pairs) — so count.
  u64 hc[3] = {n_dup, n_zero, 0};
  if (n_inv != 0) {
    for (int attempt = 0; attempt < 2; attempt++) {
      {
        ProfScope ps(ctx, KID_HEADS, n * (u64)(12 + (w ? 8 : 0)));
        k_heads<<<nblk + 1, TB, 0, st>>>(cols, L, p.use_key, key_sorted, idx_cur, w, n, nullptr, nullptr, cnt);
      }
      LAUNCH_COUNT(ctx);
      TRY(read_back(ctx, cnt, 3, hc));
      if (hc[2] == 0) break;
      // a bucket did not fit a shared-memory chunk (heavy key skew): redo with the plain LSD sequence
      if (attempt == 1) { set_error("consolidate: sort fallback failed"); return DBSP_ERR_CUDA; }
      CUDA_TRY(cudaMemsetAsync(cnt, 0, 24, st));
      TRY(sort_rows(true));
    }
  }

  MCols oc;
  i64* ow;
  if (hc[0] == 0 && hc[1] == 0) {
    Batch* b;
    TRY(batch_alloc(ctx, s, n, &b, &oc, &ow));
    {
      ProfScope ps(ctx, KID_EMIT, n * (u64)(12 + (w ? 8 : 0)) + n * (u64)(L + 1) * 8);
      k_emit_unique<<<nblk, TB, 0, st>>>(cols, p, key_sorted, idx_cur, w, n, oc, ow);
    }
    LAUNCH_COUNT(ctx);
    *out = b;
    return DBSP_OK;
  }

  // duplicates and/or zero weights: fused reduce-by-key over the sorted order
  const u32 ntiles = (u32)((n + RBK_TILE - 1) / RBK_TILE);
  BufP tbuf;
  TRY(dev_alloc(ctx, (size_t)ntiles * (sizeof(RbkTile) + sizeof(RbkCarry)) + 64, &tbuf));
  RbkTile* tiles = (RbkTile*)tbuf->p;
  RbkCarry* carry = (RbkCarry*)(tiles + ntiles);
  u32* d_total = (u32*)(carry + ntiles);
  MCols none;
  for (int l = 0; l < MAXL; l++) none.c[l] = nullptr;
  {
    ProfScope pseg(ctx, KID_SEG_REDUCE, 0);
    k_rbk<0><<<ntiles, RBK_THREADS, 0, st>>>(cols, p, key_sorted, idx_cur, w, n, tiles, nullptr, none, nullptr);
    k_rbk_scan<<<1, 1024, 0, st>>>(tiles, ntiles, carry, d_total);
    ctx->kernel_launches += 2;
  }
  u32 nout;
  TRY(read_back32(ctx, d_total, &nout));
  if (nout == 0) { *out = batch_new_empty(ctx, s); return DBSP_OK; }
  Batch* b;
  TRY(batch_alloc(ctx, s, nout, &b, &oc, &ow));
  {
    ProfScope pseg(ctx, KID_SEG_REDUCE, n * (u64)(L + 1) * 8 * 2 + (u64)nout * (L + 1) * 8);
    k_rbk<1><<<ntiles, RBK_THREADS, 0, st>>>(cols, p, key_sorted, idx_cur, w, n, tiles, carry, oc, ow);
  }
  LAUNCH_COUNT(ctx);
  *out = b;
  return DBSP_OK;
}
