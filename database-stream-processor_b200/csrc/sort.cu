// sort.cu — the sort inside Batch::from_tuples (K1): hand-written radix sort of (packed key word, row id) pairs.
//
// Replaces the comparison sort of consolidation (crates/dbsp/src/trace/consolidation/mod.rs:92-111: "90% of the
// work done while joining or merging"; consolidation/quicksort.rs:12-33).  Only the bits that the packed key word
// actually uses are sorted (consolidate.cu builds the word).  Two building blocks, both stable:
//
//  (1) k_rs_pass — one 8-bit digit pass over HBM ("onesweep"): every tile of 3072 pairs is ranked in shared memory
//      (warp-private digit counters, warp match for equal digits, no atomics), the tile's 256 digit counts are
//      published to a per-tile status array and the tile's global offsets come from a decoupled look-back over the
//      preceding tiles (one chain per digit, 256 threads), so a pass reads every pair once and writes it once; the
//      pairs are first placed in tile-sorted order in shared memory so that the global stores are contiguous runs.
//      The digit histograms of ALL passes are taken by one kernel up front (k_rs_hist_all).
//  (2) k_chunk_sort — a whole range of rows ("chunk", <= 6144 pairs) is sorted inside one CTA with shared-memory
//      digit passes, skipping every digit that is constant inside the chunk.
//
// Plan (radix_sort_pairs): an LSD sort of b bits needs ceil(b/8) trips through HBM.  Instead the top digits are
// sorted first with t HBM passes (t = 1 or 2 for the batch sizes of a step), which leaves the array partitioned into
// small buckets that are contiguous and mutually ordered; consecutive buckets are then grouped into chunks and each
// chunk is finished in shared memory: t + 1 trips instead of ceil(b/8).  Inputs whose leading lane is already
// ordered (event tables arrive in time order) need no HBM pass at all: their buckets exist already.  A bucket that
// does not fit a chunk (heavy key skew) raises a flag and the caller falls back to the plain LSD sequence.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256, RS_IPT = 12, RS_TILE = RS_THREADS * RS_IPT, RS_WARPS = RS_THREADS / 32;
constexpr int CS_THREADS = 512, CS_IPT = 12, CS_CAP = CS_THREADS * CS_IPT, CS_WARPS = CS_THREADS / 32, CS_HALF = CS_CAP / 2;
constexpr int MAX_PASSES = 8;
constexpr u64 RS_AGG = 1ull << 62, RS_PREFIX = 2ull << 62, RS_MASK = (1ull << 62) - 1;

struct PassList {
  int np;
  int lo[MAX_PASSES];
  unsigned mask[MAX_PASSES];
};

__device__ __forceinline__ u64 ldr(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void str(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// digit histograms of every pass in one read of the keys
__global__ void __launch_bounds__(256) k_rs_hist_all(const u64* __restrict__ keys, u64 n, PassList pl, u32* ghist) {
  __shared__ u32 s_h[MAX_PASSES][256];
  for (int k = threadIdx.x; k < MAX_PASSES * 256; k += 256) (&s_h[0][0])[k] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
    const u64 key = keys[i];
    for (int p = 0; p < pl.np; p++) atomicAdd(&s_h[p][(unsigned)(key >> pl.lo[p]) & pl.mask[p]], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < pl.np * 256; k += 256) {
    const u32 v = (&s_h[0][0])[k];
    if (v) atomicAdd(&ghist[k], v);
  }
}

// exclusive scan of the 256 bins of each pass (one CTA per pass)
__global__ void __launch_bounds__(256) k_rs_gscan(u32* ghist) {
  __shared__ u32 s_w[8];
  u32* h = ghist + (size_t)blockIdx.x * 256;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 v = h[tid];
  u32 incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  u32 off = 0;
  for (int k = 0; k < wid; k++) off += s_w[k];
  h[tid] = off + incl - v;
}

// Stable ranks of a warp's items among the items of the same digit that precede them in the warp's blocked range.
// s_wh[257] is this warp's private counter array (bin 256 = padding items).
template <int IPT>
__device__ __forceinline__ void warp_digit_ranks(const u64 (&key)[IPT], const bool (&valid)[IPT], int lo, unsigned mask, u32* s_wh,
                                                 unsigned short (&rank)[IPT]) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1;
#pragma unroll
  for (int r = 0; r < IPT; r++) {
    const unsigned d = valid[r] ? ((unsigned)(key[r] >> lo) & mask) : 256u;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    u32 old = 0;
    if (lane == leader) {
      old = s_wh[d];
      s_wh[d] = old + __popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = (unsigned short)(old + __popc(m & lt));
    __syncwarp();
  }
}

// One digit pass over HBM.  status: u64[ntiles * 256], zeroed; ticket: u32, zeroed; gbase: exclusive scan of the
// global digit histogram of this pass.
__global__ void __launch_bounds__(RS_THREADS, 4)
k_rs_pass(const u64* __restrict__ kin, const u32* __restrict__ iin, u64* __restrict__ kout, u32* __restrict__ iout, u64 n, int lo,
          unsigned mask, const u32* __restrict__ gbase, u64* status, u32* ticket) {
  __shared__ u64 s_key[RS_TILE];
  __shared__ u32 s_id[RS_TILE];
  __shared__ u32 s_wh[RS_WARPS][257];
  __shared__ u32 s_dstart[256];
  __shared__ long long s_gb[256];
  __shared__ u32 s_wsum[RS_WARPS];
  __shared__ u32 s_tile;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);   // tiles ordered by arrival: every predecessor of a look-back is running
  for (int k = tid; k < RS_WARPS * 257; k += RS_THREADS) (&s_wh[0][0])[k] = 0;
  __syncthreads();
  const u32 tile = s_tile;
  const u64 base = (u64)tile * RS_TILE;
  const int cnt = (int)((n - base) < (u64)RS_TILE ? (n - base) : (u64)RS_TILE);

  u64 key[RS_IPT];
  u32 id[RS_IPT];
  bool valid[RS_IPT];
  unsigned short rank[RS_IPT];
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    const int j = wid * (RS_IPT * 32) + r * 32 + lane;   // warp-blocked: item order = (warp, round, lane)
    valid[r] = j < cnt;
    key[r] = valid[r] ? kin[base + j] : 0;
    id[r] = valid[r] ? iin[base + j] : 0;
  }
  warp_digit_ranks<RS_IPT>(key, valid, lo, mask, s_wh[wid], rank);
  __syncthreads();

  // per digit: exclusive prefix over the warps, tile count, look-back over the preceding tiles
  {
    const int d = tid;   // RS_THREADS == 256 digits
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      const u32 v = s_wh[w][d];
      s_wh[w][d] = run;
      run += v;
    }
    const u32 tile_cnt = run;
    u32 incl = tile_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_wsum[wid] = incl;
    // publish this tile's digit count before anything can wait on it
    u64 prev = 0;
    if (tile == 0) {
      str(&status[d], RS_PREFIX | (u64)tile_cnt);
    } else {
      str(&status[(u64)tile * 256 + d], RS_AGG | (u64)tile_cnt);
      long long p = (long long)tile - 1;
      while (true) {
        u64 v;
        do { v = ldr(&status[(u64)p * 256 + d]); } while ((v >> 62) == 0);
        prev += v & RS_MASK;
        if ((v >> 62) == 2) break;
        p--;
      }
      str(&status[(u64)tile * 256 + d], RS_PREFIX | (prev + tile_cnt));
    }
    __syncthreads();
    u32 off = 0;
    for (int k = 0; k < wid; k++) off += s_wsum[k];
    const u32 dstart = off + incl - tile_cnt;
    s_dstart[d] = dstart;
    s_gb[d] = (long long)gbase[d] + (long long)prev - (long long)dstart;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_IPT; r++) {
    if (valid[r]) {
      const unsigned d = (unsigned)(key[r] >> lo) & mask;
      const u32 li = s_dstart[d] + s_wh[wid][d] + rank[r];
      s_key[li] = key[r];
      s_id[li] = id[r];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    const int j = k * RS_THREADS + tid;
    if (j < cnt) {
      const u64 kk = s_key[j];
      const unsigned d = (unsigned)(kk >> lo) & mask;
      const long long o = s_gb[d] + j;
      kout[o] = kk;
      iout[o] = s_id[j];
    }
  }
}

// Finish a partitioned array in shared memory.  Bucket of row i = kin[i] >> bshift (bshift >= 64: one bucket);
// buckets are contiguous and ordered.  Window w = rows [w * CS_HALF, +CS_HALF); the CTA of window w sorts the chunk
// that starts at the first bucket boundary inside its window and ends at the first bucket boundary inside the next
// window — whole buckets only, at most CS_CAP rows when no bucket exceeds CS_HALF rows (else *fail is raised).
__global__ void __launch_bounds__(CS_THREADS, 2)
k_chunk_sort(const u64* __restrict__ kin, const u32* __restrict__ iin, u64* __restrict__ kout, u32* __restrict__ iout, u64 n, int bits,
             int bshift, unsigned long long* fail) {
  extern __shared__ __align__(16) unsigned char cs_smem[];
  u64* s_key = (u64*)cs_smem;
  u32* s_id = (u32*)(s_key + CS_CAP);
  u32* s_wh = s_id + CS_CAP;   // [CS_WARPS][257]
  __shared__ u32 s_dstart[256];
  __shared__ u32 s_wsum[8];
  __shared__ unsigned long long s_first[2];
  __shared__ unsigned long long s_vary;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u64 x0 = (u64)blockIdx.x * CS_HALF;
  if (x0 >= n) return;
  if (tid == 0) { s_first[0] = ~0ull; s_first[1] = ~0ull; s_vary = 0; }
  __syncthreads();
  // first bucket boundary at or after x0 and at or after x0 + CS_HALF
#pragma unroll
  for (int side = 0; side < 2; side++) {
    const u64 from = x0 + (u64)side * CS_HALF;
    u64 best = ~0ull;
    for (int k = 0; k < CS_HALF / CS_THREADS; k++) {
      const u64 i = from + (u64)k * CS_THREADS + tid;
      if (i < n) {
        bool head = i == 0;
        if (!head && bshift < 64) head = (kin[i] >> bshift) != (kin[i - 1] >> bshift);
        if (head && i < best) best = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 y = __shfl_xor_sync(0xffffffffu, best, o);
      best = y < best ? y : best;
    }
    if (lane == 0 && best != ~0ull) atomicMin(&s_first[side], (unsigned long long)best);
  }
  __syncthreads();
  u64 s = s_first[0], e = s_first[1];
  if (s == ~0ull) {
    // no bucket starts in this window: fine at the tail of the array (the previous chunk runs to n), else a bucket
    // longer than CS_HALF rows
    if (x0 + CS_HALF < n && tid == 0) atomicOr(fail, 1ull);
    return;
  }
  if (e == ~0ull) {
    if (x0 + 2 * (u64)CS_HALF < n) { if (tid == 0) atomicOr(fail, 1ull); return; }
    e = n;
  }
  const int m = (int)(e - s);
  if (m > CS_CAP) { if (tid == 0) atomicOr(fail, 1ull); return; }

  u64 key[CS_IPT];
  u32 id[CS_IPT];
  bool valid[CS_IPT];
  unsigned short rank[CS_IPT];
  const u64 k0 = kin[s];
  u64 vary = 0;
#pragma unroll
  for (int r = 0; r < CS_IPT; r++) {
    const int j = wid * (CS_IPT * 32) + r * 32 + lane;
    valid[r] = j < m;
    key[r] = valid[r] ? kin[s + j] : k0;
    id[r] = valid[r] ? iin[s + j] : 0;
    vary |= key[r] ^ k0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vary |= __shfl_xor_sync(0xffffffffu, vary, o);
  if (lane == 0 && vary) atomicOr(&s_vary, (unsigned long long)vary);
  __syncthreads();
  vary = s_vary;
  if (bits < 64) vary &= (1ull << bits) - 1;

  for (int lo = 0; lo < bits; lo += 8) {
    const unsigned mask = (bits - lo >= 8) ? 0xffu : ((1u << (bits - lo)) - 1);
    if (((vary >> lo) & mask) == 0) continue;   // digit constant inside the chunk (block-uniform test)
    for (int k = tid; k < CS_WARPS * 257; k += CS_THREADS) s_wh[k] = 0;
    __syncthreads();
    warp_digit_ranks<CS_IPT>(key, valid, lo, mask, s_wh + wid * 257, rank);
    __syncthreads();
    if (tid < 256) {
      const int d = tid;
      u32 run = 0;
#pragma unroll
      for (int w = 0; w < CS_WARPS; w++) {
        const u32 v = s_wh[w * 257 + d];
        s_wh[w * 257 + d] = run;
        run += v;
      }
      u32 incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
      }
      if (lane == 31) s_wsum[wid] = incl;
      s_dstart[d] = incl - run;   // exclusive inside the warp of digits; the warp offsets are added below
    }
    __syncthreads();
    if (tid < 256) {
      u32 off = 0;
      for (int k = 0; k < wid; k++) off += s_wsum[k];
      s_dstart[tid] += off;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CS_IPT; r++) {
      if (valid[r]) {
        const unsigned d = (unsigned)(key[r] >> lo) & mask;
        const u32 li = s_dstart[d] + s_wh[wid * 257 + d] + rank[r];
        s_key[li] = key[r];
        s_id[li] = id[r];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CS_IPT; r++) {
      const int j = wid * (CS_IPT * 32) + r * 32 + lane;
      if (valid[r]) { key[r] = s_key[j]; id[r] = s_id[j]; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < CS_IPT; r++) {
    const int j = wid * (CS_IPT * 32) + r * 32 + lane;
    if (valid[r]) { kout[s + j] = key[r]; iout[s + j] = id[r]; }
  }
}

constexpr size_t CS_SMEM = (size_t)CS_CAP * 12 + (size_t)CS_WARPS * 257 * 4;

}  // namespace

// Sorts the n (key, id) pairs in (ka, ia) by the low `bits` bits of the key, stably.  (kb, ib) are scratch of the
// same size.  presorted_top_bits > 0: the top that many of the `bits` are already non-decreasing along the array.
// force_lsd: plain LSD sequence (the fallback after *fail was raised).  *fail (device word, may be raised by this
// call) must be checked by the caller after the stream reaches the result.  The result lands in *key_out / *idx_out.
int32_t radix_sort_pairs(Ctx* ctx, u64* ka, u64* kb, u32* ia, u32* ib, u64 n, int bits, int presorted_top_bits, bool force_lsd,
                         unsigned long long* fail, u64** key_out, u32** idx_out, int* hbm_passes) {
  cudaStream_t st = ctx->stream;
  *key_out = ka;
  *idx_out = ia;
  if (hbm_passes) *hbm_passes = 0;
  if (n <= 1 || bits <= 0) return DBSP_OK;
  if (n >= 0xffffffffull) { set_error("sort: 2^32-1 or more rows"); return DBSP_ERR_UNSUPPORTED; }
  if (bits > 64) bits = 64;

  // ---- plan -------------------------------------------------------------------------------------
  PassList pl;
  pl.np = 0;
  bool chunk = false;
  int bshift = 64;
  if (!force_lsd) {
    if (n <= (u64)CS_HALF) {
      chunk = true;               // one chunk, one bucket
    } else if (presorted_top_bits > 0) {
      chunk = true;               // the buckets exist already
      bshift = bits - (presorted_top_bits < bits ? presorted_top_bits : bits);
    } else {
      int t = 1;
      while ((n >> (8 * t)) > 512 && t < MAX_PASSES) t++;   // average bucket of <= 512 rows after t top digits
      if (8 * t + 8 < bits) {
        chunk = true;
        bshift = bits - 8 * t;
        for (int j = 0; j < t; j++) { pl.lo[pl.np] = bits - 8 * (t - j); pl.mask[pl.np] = 0xffu; pl.np++; }
      }
    }
  }
  if (!chunk) {
    for (int lo = 0; lo < bits; lo += 8) {
      pl.lo[pl.np] = lo;
      pl.mask[pl.np] = (bits - lo >= 8) ? 0xffu : ((1u << (bits - lo)) - 1);
      pl.np++;
    }
  }

  u64* kc = ka;
  u64* kn = kb;
  u32* ic = ia;
  u32* in_ = ib;
  if (pl.np > 0) {
    const u32 ntiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    BufP aux;
    // ghist[np][256] u32 | tickets[np] u32 | status[np][ntiles*256] u64
    const size_t head_bytes = ((size_t)pl.np * 256 * 4 + (size_t)pl.np * 4 + 15) & ~15ull;
    const size_t status_bytes = (size_t)pl.np * ntiles * 256 * 8;
    TRY(dev_alloc(ctx, head_bytes + status_bytes, &aux));
    CUDA_TRY(cudaMemsetAsync(aux->p, 0, head_bytes + status_bytes, st));
    u32* ghist = (u32*)aux->p;
    u32* tickets = ghist + (size_t)pl.np * 256;
    u64* status = (u64*)((char*)aux->p + head_bytes);
    {
      ProfScope ps(ctx, KID_RADIX_SORT, n * 8);
      const unsigned g = (unsigned)std::min<u64>((n + 255) / 256, (u64)ctx->sm_count * 8);
      k_rs_hist_all<<<g, 256, 0, st>>>(kc, n, pl, ghist);
      k_rs_gscan<<<pl.np, 256, 0, st>>>(ghist);
    }
    ctx->kernel_launches += 2;
    for (int p = 0; p < pl.np; p++) {
      {
        ProfScope ps(ctx, KID_RADIX_SORT, n * 24);   // every pair read once and written once
        k_rs_pass<<<ntiles, RS_THREADS, 0, st>>>(kc, ic, kn, in_, n, pl.lo[p], pl.mask[p], ghist + (size_t)p * 256,
                                                 status + (size_t)p * ntiles * 256, tickets + p);
      }
      LAUNCH_COUNT(ctx);
      std::swap(kc, kn);
      std::swap(ic, in_);
    }
    if (hbm_passes) *hbm_passes = pl.np;
  }
  if (chunk) {
    CUDA_TRY(cudaFuncSetAttribute(k_chunk_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_SMEM));
    const unsigned nwin = (unsigned)((n + CS_HALF - 1) / CS_HALF);
    {
      ProfScope ps(ctx, KID_CHUNK_SORT, n * 24);
      k_chunk_sort<<<nwin, CS_THREADS, CS_SMEM, st>>>(kc, ic, kn, in_, n, bits, bshift, fail);
    }
    LAUNCH_COUNT(ctx);
    std::swap(kc, kn);
    std::swap(ic, in_);
    if (hbm_passes) *hbm_passes += 1;
  }
  *key_out = kc;
  *idx_out = ic;
  return DBSP_OK;
}
