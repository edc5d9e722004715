// fe_device.cuh — device-side building blocks shared by the kernels of fe_kernels.cuh:
// exact (unfused) float predicates, a block-wide stable LSD radix pass, lock-free union-find,
// block scans.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fe {

#define FE_FULL 0xffffffffu

// ---- exact float arithmetic ----------------------------------------------------------------
// FLANN L2_Simple<float>: result = ((0 + dx*dx) + dy*dy) + dz*dz, every operation rounded to
// float, no contraction.  The translation unit is built with -fmad=false; the _rn intrinsics
// make the intent explicit where a comparison hangs off the result.
__device__ __forceinline__ float l2_simple(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- TMA 1-D bulk copy global -> shared with an mbarrier (sm_90+/sm_100a) -------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// dst, src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// ---- block scans -----------------------------------------------------------------------------
// Exclusive prefix of `v` over the threads of the block in thread order; *total gets the sum.
// `sm` needs NT/32 + 1 ints.  Two barriers.
template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FE_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();  // protects sm against a previous use
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NT / 32; i++) {
    int c = sm[i];
    if (i < w) woff += c;
    tot += c;
  }
  *total = tot;
  return woff + inc - v;
}

// ---- block-wide stable radix pass --------------------------------------------------------------
// One LSD pass over n items with 8-bit digits.  digit_of(i) -> 0..255 for item i of the input
// order; move(i, pos) places item i at output position pos.  Stable: equal digits keep input
// order.  wc: NT/32 * 257 counters of CntT; base: 256 + 32 uint32 (after the call base[d] is the
// END offset of digit d).  Items are processed in tiles of NT*ROUNDS; inside a tile every warp
// owns a contiguous run and ranks its items with __match_any_sync against per-warp counters.
template <int NT, typename CntT, class DigitFn, class MoveFn>
__device__ __forceinline__ void block_radix_pass(int n, DigitFn digit_of, MoveFn move, CntT* wc, unsigned* base) {
  constexpr int NW = NT / 32;
  constexpr int ROUNDS = 4;
  constexpr int TILE = NT * ROUNDS;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned* wsum = base + 256;
  __syncthreads();
  for (int i = tid; i < 256; i += NT) base[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += NT) atomicAdd(&base[digit_of(i)], 1u);
  __syncthreads();
  // exclusive scan of the 256 digit counts
  if constexpr (NT >= 256) {
    unsigned v = 0, inc = 0;
    if (tid < 256) {
      v = base[tid];
      inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(FE_FULL, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) wsum[w] = inc;
    }
    __syncthreads();
    if (tid < 256) {
      unsigned off = 0;
      for (int i = 0; i < w; i++) off += wsum[i];
      base[tid] = off + inc - v;
    }
  } else {
    constexpr int DPT = 256 / NT;  // consecutive digits per thread
    unsigned loc[DPT];
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < DPT; k++) { loc[k] = base[tid * DPT + k]; sum += loc[k]; }
    unsigned inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned t = __shfl_up_sync(FE_FULL, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned off = 0;
    for (int i = 0; i < w; i++) off += wsum[i];
    unsigned run = off + inc - sum;
#pragma unroll
    for (int k = 0; k < DPT; k++) { base[tid * DPT + k] = run; run += loc[k]; }
  }
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += TILE) {
    for (int i = tid; i < NW * 257; i += NT) wc[i] = 0;
    __syncthreads();
    int dg[ROUNDS];
    unsigned rk[ROUNDS];
    CntT* mywc = wc + w * 257;
    const int start = t0 + w * (32 * ROUNDS);
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int i = start + r * 32 + lane;
      const int d = (i < n) ? (int)digit_of(i) : 256;
      dg[r] = d;
      const unsigned peers = __match_any_sync(FE_FULL, d);
      const int leader = __ffs(peers) - 1;
      unsigned prev = 0;
      if (lane == leader) { prev = mywc[d]; mywc[d] = (CntT)(prev + __popc(peers)); }
      prev = __shfl_sync(FE_FULL, prev, leader);
      rk[r] = prev + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    for (int dd = tid; dd < 256; dd += NT) {
      unsigned run = base[dd];
      for (int ww = 0; ww < NW; ww++) {
        unsigned c = wc[ww * 257 + dd];
        wc[ww * 257 + dd] = (CntT)run;
        run += c;
      }
      base[dd] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int i = start + r * 32 + lane;
      if (i < n) move(i, (int)((unsigned)mywc[dg[r]] + rk[r]));
    }
    __syncthreads();
  }
}

// ---- lock-free union-find (root = smallest index of the component) -------------------------------
__device__ __forceinline__ unsigned uf_find(unsigned* parent_, unsigned a) {
  volatile unsigned* parent = parent_;  // other threads link roots concurrently
  unsigned p = parent[a];
  while (p != a) {
    unsigned gp = parent[p];
    if (gp != p) parent[a] = gp;  // path halving; parents only ever move towards the root
    a = p;
    p = gp;
  }
  return a;
}

// Read-only find for the flatten pass: with every thread storing its own final label, a
// concurrent path-halving store could overwrite an already flattened entry with a stale ancestor.
__device__ __forceinline__ unsigned uf_find_readonly(const unsigned* parent_, unsigned a) {
  const volatile unsigned* parent = parent_;
  unsigned p = parent[a];
  while (p != a) { a = p; p = parent[a]; }
  return a;
}

__device__ __forceinline__ void uf_union(unsigned* parent, unsigned a, unsigned b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { unsigned t = a; a = b; b = t; }
    const unsigned old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

__device__ __forceinline__ int bits_for(int v) {  // bits needed to represent values 0..v
  return v <= 0 ? 0 : 32 - __clz(v);
}

}  // namespace fe
