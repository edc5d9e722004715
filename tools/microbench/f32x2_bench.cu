// Microbenchmark: the unfused squared-distance test (FLANN L2_Simple) for two query points per thread,
// scalar FP32 vs Blackwell packed f32x2 (FADD2 / FFMA2 with an opaque -0 addend so that ptxas cannot
// contract mul+add — it does contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 sub2(float2 a, float b) { float2 r;
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { float2 r;
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
__device__ __forceinline__ float2 sq2(float2 a, float nz) { float2 r;
  asm volatile("{.reg .b64 ra, rz, rc; mov.b64 ra, {%2,%3}; mov.b64 rz, {%4,%4}; fma.rn.f32x2 rc, ra, ra, rz; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(nz)); return r; }
template <bool PACKED>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ pts, const float4* __restrict__ q, int m, int* out, float r2, float nz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float4 p1 = pts[2 * i], p2 = pts[2 * i + 1];
  int c1 = 0, c2 = 0;
  for (int j = 0; j < m; j++) {
    const float4 c = q[j];
    if (PACKED) {
      float2 s = sq2(sub2(make_float2(p1.x, p2.x), c.x), nz);
      s = add2(s, sq2(sub2(make_float2(p1.y, p2.y), c.y), nz));
      s = add2(s, sq2(sub2(make_float2(p1.z, p2.z), c.z), nz));
      c1 += s.x < r2; c2 += s.y < r2;
    } else {
      float dx = __fsub_rn(p1.x, c.x), dy = __fsub_rn(p1.y, c.y), dz = __fsub_rn(p1.z, c.z);
      float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      c1 += s < r2;
      dx = __fsub_rn(p2.x, c.x); dy = __fsub_rn(p2.y, c.y); dz = __fsub_rn(p2.z, c.z);
      s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      c2 += s < r2;
    }
  }
  out[2 * i] = c1; out[2 * i + 1] = c2;
}
int main() {
  const int nthr = 148 * 8 * 256, m = 4096;
  float4 *pts, *q; int *o1, *o2;
  cudaMalloc(&pts, sizeof(float4) * 2 * nthr); cudaMalloc(&q, sizeof(float4) * m); cudaMalloc(&o1, 8 * nthr); cudaMalloc(&o2, 8 * nthr);
  float4* h = (float4*)malloc(sizeof(float4) * 2 * nthr);
  unsigned s = 1; auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  for (int i = 0; i < 2 * nthr; i++) h[i] = make_float4(rnd(), rnd(), rnd(), 0);
  cudaMemcpy(pts, h, sizeof(float4) * 2 * nthr, cudaMemcpyHostToDevice); cudaMemcpy(q, h, sizeof(float4) * m, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; rep++) {
    float ms1, ms2;
    cudaEventRecord(a); k<false><<<nthr / 256, 256>>>(pts, q, m, o1, 0.25f, -0.0f); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms1, a, b);
    cudaEventRecord(a); k<true><<<nthr / 256, 256>>>(pts, q, m, o2, 0.25f, -0.0f); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms2, a, b);
    const double tests = 2.0 * nthr * m;
    printf("scalar %.3f ms (%.2f Ttests/s)   packed %.3f ms (%.2f Ttests/s)\n", ms1, tests / ms1 / 1e9, ms2, tests / ms2 / 1e9);
  }
  int* r1 = (int*)malloc(8 * nthr); int* r2 = (int*)malloc(8 * nthr);
  cudaMemcpy(r1, o1, 8 * nthr, cudaMemcpyDeviceToHost); cudaMemcpy(r2, o2, 8 * nthr, cudaMemcpyDeviceToHost);
  long long diff = 0; for (int i = 0; i < 2 * nthr; i++) diff += r1[i] != r2[i];
  printf("counts differing between scalar and packed: %lld of %d (%s)\n", diff, 2 * nthr, cudaGetErrorString(cudaGetLastError()));
  return diff != 0;
}
