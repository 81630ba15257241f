// ctx.cu — context, stream-ordered memory, small host<->device plumbing.
#include <cub/device/device_scan.cuh>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

static inline double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
void print_host_stats(Ctx* c) {
  if (getenv("DBSP_HOST_STATS"))
    fprintf(stderr, "[dbsp host] allocs %llu in %.1f ms; read-back syncs %llu waited %.1f ms; launches %llu\n",
            (unsigned long long)c->n_alloc, c->t_alloc_us / 1e3, (unsigned long long)c->n_sync, c->t_sync_us / 1e3,
            (unsigned long long)c->kernel_launches);
}
const char* get_error() { return g_err.c_str(); }

static const char* KNAMES[KID_COUNT] = {
    "merge_tiles", "merge_partition", "probe_ranges", "probe_fill", "project_rows", "radix_sort", "pack_keys",
    "heads", "emit", "minmax", "seg_reduce", "lookup", "compact", "scan", "agg_pick", "misc"};
const char* kernel_name(int id) { return (id >= 0 && id < KID_COUNT) ? KNAMES[id] : "?"; }

ProfScope::ProfScope(Ctx* ctx, int id, u64 bytes) : c(ctx) {
  if (!ctx->prof_on) return;
  ProfRec r;
  r.id = id;
  r.bytes = bytes;
  for (cudaEvent_t* e : {&r.a, &r.b}) {
    if (!ctx->ev_pool.empty()) { *e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
    else cudaEventCreate(e);
  }
  cudaEventRecord(r.a, ctx->stream);
  idx = (long)ctx->prof.size();
  ctx->prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(c->prof[idx].b, c->stream);
}

// size classes: 8 per doubling above 4 KiB (<= 12.5% slack), 512-byte steps below
static size_t size_class(size_t b) {
  if (b <= 4096) return (b + 511) & ~(size_t)511;
  size_t p = (size_t)1 << (63 - __builtin_clzll((unsigned long long)b));
  size_t step = p >> 3;
  return (b + step - 1) / step * step;
}

// Blocks are carved from large slabs with a bump pointer (one cudaMalloc per
// slab: a cudaMalloc per growing trace batch cost milliseconds each) and
// recycled through per-class free lists.
static const size_t SLAB_BYTES = (size_t)1 << 31;   // 2 GiB

void pool_release_all(Ctx* ctx) {
  for (void* p : ctx->slabs) cudaFree(p);
  ctx->slabs.clear();
  ctx->free_blocks.clear();
  ctx->slab_cur = nullptr;
  ctx->slab_left = 0;
  ctx->pool_reserved = ctx->pool_cached = 0;
}

DevBuf::~DevBuf() {
  if (!ctx) return;
  if (p && !ctx->destroyed) {   // after ctx_destroy the slabs are already gone
    ctx->free_blocks[cls].push_back(p);
    ctx->pool_cached += cls;
  }
  if (ctx->live_bufs.fetch_sub(1) == 1 && ctx->destroyed) delete ctx;
}

int32_t dev_alloc(Ctx* ctx, size_t bytes, BufP* out) {
  auto b = std::make_shared<DevBuf>();
  b->ctx = ctx;
  ctx->live_bufs.fetch_add(1);
  b->bytes = bytes;
  if (bytes == 0) bytes = 16;
  b->cls = size_class(bytes);
  double t0 = now_us();
  auto it = ctx->free_blocks.find(b->cls);
  if (it != ctx->free_blocks.end() && !it->second.empty()) {
    b->p = it->second.back();
    it->second.pop_back();
    ctx->pool_cached -= b->cls;
  } else {
    size_t need = (b->cls + 255) & ~(size_t)255;
    if (need > ctx->slab_left) {
      size_t slab = need > SLAB_BYTES ? need : SLAB_BYTES;
      void* p = nullptr;
      cudaError_t e = cudaMalloc(&p, slab);
      if (e != cudaSuccess && slab > need) {   // not enough room for a full slab: take just what is needed
        cudaGetLastError();
        slab = need;
        e = cudaMalloc(&p, slab);
      }
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string("cudaMalloc(") + std::to_string(slab) + "): " + cudaGetErrorString(e) +
                  " (pool reserved " + std::to_string(ctx->pool_reserved) + " B, cached " + std::to_string(ctx->pool_cached) + " B)");
        return DBSP_ERR_CUDA;
      }
      ctx->slabs.push_back(p);
      ctx->slab_cur = (char*)p;
      ctx->slab_left = slab;
      ctx->pool_reserved += slab;
    }
    b->p = ctx->slab_cur;
    ctx->slab_cur += need;
    ctx->slab_left -= need;
  }
  ctx->t_alloc_us += now_us() - t0;
  ctx->n_alloc++;
  *out = b;
  return DBSP_OK;
}

int32_t batch_alloc(Ctx* ctx, const dbsp_schema& s, u64 n, Batch** out, MCols* cols, i64** w) {
  int L = s.n_key_lanes + s.n_val_lanes;
  u64 cap = (n + 31) & ~31ull;   // keep every lane 256-byte aligned
  if (cap == 0) cap = 32;
  BufP buf;
  TRY(dev_alloc(ctx, (size_t)cap * 8 * (L + 1), &buf));
  Batch* b = new Batch();
  b->s = s;
  b->n = n;
  b->ctx = ctx;
  b->bufs.push_back(buf);
  u64* base = (u64*)buf->p;
  for (int l = 0; l < L; l++) {
    b->col[l] = base + (size_t)l * cap;
    if (cols) cols->c[l] = base + (size_t)l * cap;
  }
  b->w = (const i64*)(base + (size_t)L * cap);
  if (w) *w = (i64*)(base + (size_t)L * cap);
  *out = b;
  return DBSP_OK;
}

Batch* batch_new_empty(Ctx* ctx, const dbsp_schema& s) {
  Batch* b = new Batch();
  b->s = s;
  b->n = 0;
  b->ctx = ctx;
  b->nkeys = 0;
  return b;
}

void batch_unref(Batch* b) {
  if (!b) return;
  if (b->refs.fetch_sub(1) == 1) delete b;
}

int32_t read_back(Ctx* ctx, const void* dsrc, size_t count_u64, u64* hdst) {
  if (count_u64 > 256) { set_error("read_back: too large"); return DBSP_ERR_INVALID; }
  double t0 = now_us();
  CUDA_TRY(cudaMemcpyAsync(ctx->h_scratch, dsrc, count_u64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->t_sync_us += now_us() - t0;
  ctx->n_sync++;
  for (size_t i = 0; i < count_u64; i++) hdst[i] = ctx->h_scratch[i];
  ctx->d2h_bytes += count_u64 * 8;
  return DBSP_OK;
}

int32_t read_back32(Ctx* ctx, const void* dsrc, u32* hdst) {
  double t0 = now_us();
  CUDA_TRY(cudaMemcpyAsync(ctx->h_scratch, dsrc, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->t_sync_us += now_us() - t0;
  ctx->n_sync++;
  *hdst = *(u32*)ctx->h_scratch;
  ctx->d2h_bytes += 4;
  return DBSP_OK;
}

// Device-wide prefix sums are plumbing between the hand-written kernels (the
// hot merge kernel carries its own decoupled look-back instead).
int32_t exclusive_scan_u32(Ctx* ctx, const u32* in, u32* out, u64 n) {
  // out has n+1 entries: out[n] = total.  Scan n+1 inputs where in[n] is
  // ignored by construction: callers allocate in with n+1 entries.
  size_t tmp_bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (size_t)(n + 1), ctx->stream));
  BufP tmp;
  TRY(dev_alloc(ctx, tmp_bytes, &tmp));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp->p, tmp_bytes, in, out, (size_t)(n + 1), ctx->stream));
  LAUNCH_COUNT(ctx);
  return DBSP_OK;
}

int32_t inclusive_scan_i64(Ctx* ctx, const i64* in, i64* out, u64 n) {
  size_t tmp_bytes = 0;
  CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, in, out, (size_t)n, ctx->stream));
  BufP tmp;
  TRY(dev_alloc(ctx, tmp_bytes, &tmp));
  CUDA_TRY(cub::DeviceScan::InclusiveSum(tmp->p, tmp_bytes, in, out, (size_t)n, ctx->stream));
  LAUNCH_COUNT(ctx);
  return DBSP_OK;
}
