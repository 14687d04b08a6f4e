// Microbenchmarks behind the conv1 design: (1) tcgen05.ld (TMEM -> registers) bandwidth per SM with 4 / 8 / 16
// reader warps, (2) legacy mma.sync rates (m16n8k16 f16->f32, m16n8k32 s8->s32) per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_hmma tmem_hmma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__global__ void k_tmem_ld(int iters, long long *cyc, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = ((it + (warp >> 2) * 4) & 7) * 64;
    if constexpr (X == 16) {
      uint32_t r[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(base + col + k * 16) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= r[i];
      }
    } else {
      uint32_t r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(base + col + k * 8) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= r[i];
      }
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

// 8 independent accumulator tiles per warp
template <int KIND>
__global__ void k_mma_sync(int iters, long long *cyc, float *sink) {
  float c[8][4];
  int ci[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j] = 0.f; ci[i][j] = 0; }
  uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x55, b1 = a0 ^ 0x33;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (KIND == 0) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else {
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(ci[i][0]), "+r"(ci[i][1]), "+r"(ci[i][2]), "+r"(ci[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j] + (float)ci[i][j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  long long *cyc; uint32_t *sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int x : {16, 8}) {
      if (x == 16) { k_tmem_ld<16><<<148, warps * 32>>>(iters, cyc, sink); k_tmem_ld<16><<<148, warps * 32>>>(iters, cyc, sink); }
      else { k_tmem_ld<8><<<148, warps * 32>>>(iters, cyc, sink); k_tmem_ld<8><<<148, warps * 32>>>(iters, cyc, sink); }
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * 64 * 4 * 32 * warps;   // 64 columns x 4 B x 32 lanes per warp-iteration
      printf("tcgen05.ld x%-2d %2d warps: %8.1f cycles/iter  -> %6.1f B/cycle/SM (%s)\n", x, warps, (double)c / iters, bytes / c, cudaGetErrorString(e));
    }
  }
  for (int kind : {0, 1}) {
    for (int warps : {4, 8, 16}) {
      if (kind == 0) { k_mma_sync<0><<<148, warps * 32>>>(iters, cyc, (float *)sink); k_mma_sync<0><<<148, warps * 32>>>(iters, cyc, (float *)sink); }
      else { k_mma_sync<1><<<148, warps * 32>>>(iters, cyc, (float *)sink); k_mma_sync<1><<<148, warps * 32>>>(iters, cyc, (float *)sink); }
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double macs = (double)iters * 8 * warps * (kind == 0 ? 16 * 8 * 16 : 16 * 8 * 32);
      printf("mma.sync %s %2d warps: %7.2f cycles per mma per SMSP -> %7.1f MAC/cycle/SM (%s)\n", kind == 0 ? "m16n8k16 f16->f32" : "m16n8k32 s8->s32 ",
             warps, (double)c / iters / 8 / (warps / 4.0), macs / c, cudaGetErrorString(e));
    }
  }
  return 0;
}
