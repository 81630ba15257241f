// merge_calibration.cu — what can the merge kernel's memory pattern reach?
//
// Stand-alone calibration (not part of the library): the same HBM streams as
// k_merge_tiles<2> on 2 x N (u64 key, u64 val, i64 weight) rows — 4 lane arrays
// staged per tile with TMA bulk copies, 2 weight arrays read per thread, 3
// output arrays — but no merge logic at all (tile t takes rows [t*H,(t+1)*H) of
// A and of B and writes them back to back).  Variants:
//   0  flat grid-stride copy, 16-byte loads/stores          (the copy ceiling for 6 in / 3 out streams)
//   1  tile: TMA stage lanes -> smem -> 8-byte coalesced stores; weights LDG -> regs -> per-thread runs
//   2  like 1, weights stored coalesced (through the same loop as the lanes)
//   3  like 1, plus one decoupled-look-back style status publish/poll per tile (32-wide)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o merge_calibration merge_calibration.cu
// Run:    ./merge_calibration [rows_per_input=50000000]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;

constexpr int THREADS = 256, IPT = 9, TILE = THREADS * IPT, H = TILE / 2, S = TILE + 8;

__device__ __forceinline__ void tma_g2s(void* sdst, const void* gsrc, unsigned bytes, u64* mbar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(sdst), m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc),
               "r"(bytes), "r"(m)
               : "memory");
}
__device__ __forceinline__ void mbar_init(u64* mbar, unsigned c) {
  unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(c) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(u64* mbar, unsigned bytes) {
  unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* mbar, unsigned parity) {
  unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}" ::"r"(m),
      "r"(parity)
      : "memory");
}

struct Arr { const u64 *a0, *a1, *wa, *b0, *b1, *wb; u64 *o0, *o1, *wo; };

__global__ void k_flat(Arr p, u64 n) {   // n rows per input; 16-byte accesses
  const ulonglong2 *a0 = (const ulonglong2*)p.a0, *a1 = (const ulonglong2*)p.a1, *wa = (const ulonglong2*)p.wa;
  const ulonglong2 *b0 = (const ulonglong2*)p.b0, *b1 = (const ulonglong2*)p.b1, *wb = (const ulonglong2*)p.wb;
  ulonglong2 *o0 = (ulonglong2*)p.o0, *o1 = (ulonglong2*)p.o1, *wo = (ulonglong2*)p.wo;
  u64 h = n / 2;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < h; i += (u64)gridDim.x * blockDim.x) {
    o0[i] = a0[i]; o1[i] = a1[i]; wo[i] = wa[i];
    o0[h + i] = b0[i]; o1[h + i] = b1[i]; wo[h + i] = wb[i];
  }
}

template <int VAR>
__global__ void __launch_bounds__(THREADS, 5) k_tile(Arr p, u64 n, u64* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64* sl = (u64*)smem;   // 2 lanes x S slots: A rows at [0,H), B rows at [H,2H)
  __shared__ __align__(8) u64 s_mbar;
  __shared__ u64 s_base;
  const int tid = threadIdx.x;
  const u64 t = blockIdx.x;
  const u64 r0 = t * H;   // first row of both inputs
  if (tid == 0) mbar_init(&s_mbar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect(&s_mbar, 4u * H * 8u);
    tma_g2s(sl, p.a0 + r0, H * 8, &s_mbar);
    tma_g2s(sl + S, p.a1 + r0, H * 8, &s_mbar);
    tma_g2s(sl + H, p.b0 + r0, H * 8, &s_mbar);
    tma_g2s(sl + S + H, p.b1 + r0, H * 8, &s_mbar);
  }
  // weights: each thread owns IPT consecutive merged positions; here the first half of the tile is A, the second B
  u64 w[IPT];
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    const int pos = tid * IPT + k;
    w[k] = pos < H ? p.wa[r0 + pos] : p.wb[r0 + pos - H];
  }
  mbar_wait(&s_mbar, 0);
  __syncthreads();
  u64 base = t * TILE;
  if (VAR == 3) {   // publish an aggregate, poll the 32 predecessors, publish a prefix (no dependence on values)
    if (t == 0) { if (tid == 0) { atomicExch(&status[0], (2ull << 62) | TILE); s_base = 0; } }
    else {
      if (tid == 0) atomicExch(&status[t], (1ull << 62) | TILE);
      if (tid < 32) {
        u64 acc = 0;
        long long q0 = (long long)t - 1;
        while (true) {
          long long q = q0 - tid;
          u64 v = 2ull << 62;
          if (q >= 0) do { v = *(volatile u64*)&status[q]; } while ((v >> 62) == 0);
          unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
          int first = pm ? __ffs(pm) - 1 : 32;
          u64 x = tid <= first ? (v & ((1ull << 62) - 1)) : 0;
          for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
          acc += x;
          if (pm) break;
          q0 -= 32;
        }
        if (tid == 0) { atomicExch(&status[t], (2ull << 62) | (acc + TILE)); s_base = acc; }
      }
    }
    __syncthreads();
    base = s_base;
  }
  for (int o = tid; o < TILE; o += THREADS) {
    p.o0[base + o] = sl[o];
    p.o1[base + o] = sl[S + o];
  }
  if (VAR == 2) {
    // coalesced weight stores need the weights in shared memory: reuse lane 0 after a barrier
    __syncthreads();
#pragma unroll
    for (int k = 0; k < IPT; k++) sl[tid * IPT + k] = w[k];
    __syncthreads();
    for (int o = tid; o < TILE; o += THREADS) p.wo[base + o] = sl[o];
  } else {
#pragma unroll
    for (int k = 0; k < IPT; k++) p.wo[base + tid * IPT + k] = w[k];
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 50000000ull;
  n = n / H * H;   // whole tiles
  Arr p;
  u64** in[6] = {(u64**)&p.a0, (u64**)&p.a1, (u64**)&p.wa, (u64**)&p.b0, (u64**)&p.b1, (u64**)&p.wb};
  for (auto q : in) { CK(cudaMalloc(q, n * 8)); CK(cudaMemset(*q, 1, n * 8)); }
  CK(cudaMalloc(&p.o0, 2 * n * 8)); CK(cudaMalloc(&p.o1, 2 * n * 8)); CK(cudaMalloc(&p.wo, 2 * n * 8));
  u64 ntiles = n / H;
  u64* status; CK(cudaMalloc(&status, ntiles * 8));
  size_t smem = (size_t)S * 2 * 8;
  CK(cudaFuncSetAttribute(k_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tile<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = (double)n * 2 * 3 * 8 * 2;   // read 6 arrays of n, write 3 arrays of 2n
  const char* names[4] = {"flat 16-byte copy", "tile: TMA lanes, per-thread weight runs", "tile: TMA lanes, coalesced weights", "tile + 32-wide look-back"};
  for (int var = 0; var < 4; var++) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
      CK(cudaMemsetAsync(status, 0, ntiles * 8));
      CK(cudaEventRecord(e0));
      if (var == 0) k_flat<<<148 * 16, 256>>>(p, n);
      else if (var == 1) k_tile<1><<<(unsigned)ntiles, THREADS, smem>>>(p, n, status);
      else if (var == 2) k_tile<2><<<(unsigned)ntiles, THREADS, smem>>>(p, n, status);
      else k_tile<3><<<(unsigned)ntiles, THREADS, smem>>>(p, n, status);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("variant %d (%s): %.3f ms  %.1f GB/s\n", var, names[var], best, bytes / best / 1e6);
  }
  return 0;
}
