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
//      in (consolidation/mod.rs:101-104) is an LSD radix sort over bit-packed
//      composite keys: the lanes' significant bits are concatenated (order
//      preserving, injective) into as few 64-bit words as possible and only
//      those bits are sorted, as (key word, row id) pairs.  The Nexmark
//      schemas pack into one word of 30-60 bits.
//  (4) Epilogue: runs of equal rows are summed with a prefix-sum difference,
//      zero sums dropped; the duplicate-free case is one unpack/gather pass.
#include <cub/device/device_radix_sort.cuh>

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
// (row i-1 > row i), mm[2L+1] = # adjacent duplicates, mm[2L+2] = # zero weights.
__global__ void k_props(Cols cols, Flips f, int L, const i64* w, u64 n, u64* mm) {
  __shared__ u64 smin[MAXL], smax[MAXL];
  __shared__ unsigned s_cnt[3];
  if (threadIdx.x < MAXL) { smin[threadIdx.x] = ~0ull; smax[threadIdx.x] = 0; }
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  u64 lmin[MAXL], lmax[MAXL];
  for (int l = 0; l < L; l++) { lmin[l] = ~0ull; lmax[l] = 0; }
  unsigned inv = 0, dup = 0, zero = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    int c = 0;   // cmp(row i-1, row i)
    for (int l = 0; l < L; l++) {
      u64 v = cols.c[l][i] ^ f.f[l];
      lmin[l] = min(lmin[l], v);
      lmax[l] = max(lmax[l], v);
      if (i > 0 && c == 0) {
        u64 pv = cols.c[l][i - 1] ^ f.f[l];
        if (pv != v) c = pv < v ? -1 : 1;
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
  }
  if ((threadIdx.x & 31) == 0) {
    if (inv) atomicAdd(&s_cnt[0], inv);
    if (dup) atomicAdd(&s_cnt[1], dup);
    if (zero) atomicAdd(&s_cnt[2], zero);
  }
  __syncthreads();
  if (threadIdx.x < L) {
    atomicMin((unsigned long long*)&mm[threadIdx.x], (unsigned long long)smin[threadIdx.x]);
    atomicMax((unsigned long long*)&mm[L + threadIdx.x], (unsigned long long)smax[threadIdx.x]);
  }
  if (threadIdx.x < 3 && s_cnt[threadIdx.x])
    atomicAdd((unsigned long long*)&mm[2 * L + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void k_init_props(u64* mm, int L) {
  int t = threadIdx.x;
  if (t < L) mm[t] = ~0ull;
  else if (t < 2 * L + 3) mm[t] = 0;
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

__global__ void k_seg_start(const u32* flags, const u32* exscan, u64 n, u32* segstart, u32 nseg) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) segstart[exscan[i]] = (u32)i;
  if (i == 0) segstart[nseg] = (u32)n;
}

__global__ void k_seg_sum(const i64* P, const u32* segstart, u32 nseg, u32* keep, i64* sums) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    u32 a = segstart[s], b = segstart[s + 1];
    i64 sum = (i64)((u64)P[b - 1] - (a ? (u64)P[a - 1] : 0ull));
    sums[s] = sum;
    keep[s] = sum != 0 ? 1u : 0u;
  } else if (s == nseg) {
    keep[nseg] = 0;
  }
}

__global__ void k_emit_seg(Cols cols, Plan p, const u64* key, const u32* idx, const u32* segstart, const u32* keep,
                           const u32* pos, const i64* sums, u32 nseg, MCols out, i64* out_w) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg || !keep[s]) return;
  u32 a = segstart[s], o = pos[s];
  if (p.use_key) {
    u64 k = key[a];
    for (int l = 0; l < p.L; l++) out.c[l][o] = unpack_lane(p, l, k);
  } else {
    u64 r = idx ? idx[a] : a;
    for (int l = 0; l < p.L; l++) out.c[l][o] = cols.c[l][r];
  }
  out_w[o] = sums[s];
}

inline int bits_for(u64 range) { return range == 0 ? 0 : 64 - __builtin_clzll(range); }

}  // namespace

int32_t consolidate_rows(Ctx* ctx, const dbsp_schema& s, const Cols& cols, const i64* w, u64 n, const BufP* adopt,
                         Batch** out) {
  const int L = s.n_key_lanes + s.n_val_lanes;
  if (n == 0) { *out = batch_new_empty(ctx, s); return DBSP_OK; }
  if (n >= (1ull << 32)) { set_error("consolidate: more than 2^32-1 rows in one batch"); return DBSP_ERR_UNSUPPORTED; }
  cudaStream_t st = ctx->stream;
  const int TB = 256;
  const unsigned nblk = (unsigned)((n + TB - 1) / TB);

  Flips f;
  for (int l = 0; l < MAXL; l++) f.f[l] = (l < L && s.lane_types[l] == DBSP_I64) ? 0x8000000000000000ull : 0;

  // ---- (1) lane ranges + order / duplicate / zero-weight census ---------------
  u64 mm[2 * MAXL + 3];
  {
    u64* dmm = ctx->d_scratch + 64;
    k_init_props<<<1, 32, 0, st>>>(dmm, L);
    int g = (int)std::min<u64>((n + TB - 1) / TB, (u64)ctx->sm_count * 8);
    {
      ProfScope ps(ctx, KID_MINMAX, n * (u64)(L + (w ? 1 : 0)) * 8);
      k_props<<<g, TB, 0, st>>>(cols, f, L, w, n, dmm);
    }
    ctx->kernel_launches += 2;
    TRY(read_back(ctx, dmm, 2 * L + 3, mm));
  }
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

    // sort (key word, row id) pairs, least significant word first
    TRY(dev_alloc(ctx, (size_t)n * 8 * 2, &kbuf));
    TRY(dev_alloc(ctx, (size_t)n * 4 * 2, &ibuf));
    u64* ka = (u64*)kbuf->p;
    u64* kb = ka + n;
    u32* ia = (u32*)ibuf->p;
    u32* ib = ia + n;
    size_t tmp_bytes = 0;
    {
      cub::DoubleBuffer<u64> dk(ka, kb);
      cub::DoubleBuffer<u32> di(ia, ib);
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, di, (int)n, 0, 64, st));
    }
    TRY(dev_alloc(ctx, tmp_bytes, &tmp));
    key_sorted = ka;
    for (int wd = 0; wd < p.W; wd++) {
      // Keys of the previous word are dead: always pack into ka.  The row ids
      // ping-pong between ia and ib.
      {
        ProfScope ps(ctx, KID_PACK, n * (u64)L * 8 + n * 12);
        if (wd == 0) {
          k_pack<<<nblk, TB, 0, st>>>(cols, p, wd, nullptr, n, ka, ia);
          idx_cur = ia;
        } else {
          k_pack<<<nblk, TB, 0, st>>>(cols, p, wd, idx_cur, n, ka, nullptr);
        }
      }
      LAUNCH_COUNT(ctx);
      key_sorted = ka;
      if (p.wbits[wd] > 0 && n > 1) {
        cub::DoubleBuffer<u64> dk(ka, kb);
        cub::DoubleBuffer<u32> di(idx_cur, idx_cur == ia ? ib : ia);
        {
          // lower bound: the (key,id) pairs read once and written once; the LSD
          // sort makes ceil(bits/8) such round trips
          ProfScope ps(ctx, KID_RADIX_SORT, n * 12 * 2);
          CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp->p, tmp_bytes, dk, di, (int)n, 0, (int)p.wbits[wd], st));
        }
        ctx->kernel_launches += (p.wbits[wd] + 7) / 8 + 1;
        idx_cur = di.Current();
        key_sorted = dk.Current();
      }
    }
  }

  // ---- (4) epilogue -------------------------------------------------------------
  // After a sort, duplicates are possible iff the input had any equal pair at
  // all — unknown from the census (it only saw adjacent pairs) — so count.
  u64 hc[2] = {n_dup, n_zero};
  if (n_inv != 0) {
    u64* cnt = ctx->d_scratch + 32;
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 16, st));
    {
      ProfScope ps(ctx, KID_HEADS, n * (u64)(12 + (w ? 8 : 0)));
      k_heads<<<nblk + 1, TB, 0, st>>>(cols, L, p.use_key, key_sorted, idx_cur, w, n, nullptr, nullptr, cnt);
    }
    LAUNCH_COUNT(ctx);
    TRY(read_back(ctx, cnt, 2, hc));
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

  // duplicates and/or zero weights: segmented sum over runs of equal rows
  ProfScope pseg(ctx, KID_SEG_REDUCE, n * (u64)(L + 1) * 8 * 2);
  BufP fbuf, wbuf, sbuf;
  TRY(dev_alloc(ctx, (size_t)(n + 1) * 4 * 2, &fbuf));
  TRY(dev_alloc(ctx, (size_t)n * 8 * 2, &wbuf));
  u32* flags = (u32*)fbuf->p;
  u32* exscan = flags + (n + 1);
  i64* ws = (i64*)wbuf->p;
  i64* P = ws + n;
  k_heads<<<nblk + 1, TB, 0, st>>>(cols, L, p.use_key, key_sorted, idx_cur, w, n, flags, ws, nullptr);
  LAUNCH_COUNT(ctx);
  TRY(exclusive_scan_u32(ctx, flags, exscan, n));
  TRY(inclusive_scan_i64(ctx, ws, P, n));
  u32 nseg;
  TRY(read_back32(ctx, exscan + n, &nseg));
  TRY(dev_alloc(ctx, (size_t)(nseg + 1) * (4 + 4 + 4 + 8) + 64, &sbuf));
  i64* sums = (i64*)sbuf->p;
  u32* segstart = (u32*)(sums + (nseg + 1));
  u32* keep = segstart + (nseg + 1);
  u32* pos = keep + (nseg + 1);
  k_seg_start<<<nblk, TB, 0, st>>>(flags, exscan, n, segstart, nseg);
  unsigned sblk = (nseg + 1 + TB - 1) / TB;
  k_seg_sum<<<sblk, TB, 0, st>>>(P, segstart, nseg, keep, sums);
  ctx->kernel_launches += 2;
  TRY(exclusive_scan_u32(ctx, keep, pos, nseg));
  u32 nout;
  TRY(read_back32(ctx, pos + nseg, &nout));
  if (nout == 0) { *out = batch_new_empty(ctx, s); return DBSP_OK; }
  Batch* b;
  TRY(batch_alloc(ctx, s, nout, &b, &oc, &ow));
  k_emit_seg<<<sblk, TB, 0, st>>>(cols, p, key_sorted, idx_cur, segstart, keep, pos, sums, nseg, oc, ow);
  LAUNCH_COUNT(ctx);
  *out = b;
  return DBSP_OK;
}
