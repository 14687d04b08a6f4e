// Microbenchmark: issue / FMA sub-pipe / ALU-pipe rates of the instruction mixes the LIF epilogue can use on
// sm_100a.  One CTA per SM, W warps per SM sub-partition; every variant runs R dependent rounds over 16 independent
// register chains and reports cycles per warp-instruction-slot.  Build: nvcc -arch=sm_100a -O3 -o fma_pipes fma_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fset(float a) { float d; asm volatile("set.ge.f32.f32 %0, %1, 0f3F800000;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ float fsel(float a) { float d; asm volatile("{.reg .pred p; setp.ge.f32 p, %1, 0f3F800000; selp.f32 %0, 0f00000000, %1, p;}" : "=f"(d) : "f"(a)); return d; }

constexpr int R = 2000;

// mode: 0 scalar FFMA x16 | 1 FFMA2 x8 (same lane-ops) | 2 8 scalar + 4 packed (same lane-ops) | 3 FSET.BF x16
//       4 setp+selp x16 | 5 16 FFMA + 8 FSET | 6 8 FFMA2 + 8 FSET | 7 8 scalar + 4 packed + 8 FSET
//       8 LIF scalar-exact (fma,sub,fma,fset,fma) x16 neurons  | 9 same fully packed | 10 same half/half
template <int MODE>
__global__ void __launch_bounds__(640, 1) k(float *out, long long *cyc, float seed) {
  float x[16];
  uint64_t p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (i + 1 + threadIdx.x);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack2(x[2 * i], x[2 * i + 1]);
  const float a = seed, b = seed * 0.5f;
  const uint64_t a2 = pack2(a, a), b2 = pack2(b, b), h2 = pack2(0.5f, 0.5f);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < R; ++r) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fma1(x[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], a2, b2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { x[2 * i] = fma1(x[2 * i], a, b); p[4 + i] = fma2(p[4 + i], a2, b2); x[2 * i + 1] = fma1(x[2 * i + 1], a, b); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fset(x[i]);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fsel(x[i]);
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[2 * i] = fma1(x[2 * i], a, b); x[2 * i + 1] = fma1(x[2 * i + 1], a, b); x[(i + 5) & 15] = fset(x[(i + 5) & 15]); }
    } else if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], a2, b2); x[i] = fset(x[i]); }
    } else if (MODE == 7) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { x[2 * i] = fma1(x[2 * i], a, b); p[4 + i] = fma2(p[4 + i], a2, b2); x[2 * i + 1] = fma1(x[2 * i + 1], a, b);
        x[8 + i] = fset(x[8 + i]); x[12 + i] = fset(x[12 + i]); }
    } else if (MODE == 8) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = fma1((float)r, a, b);
        float d; asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(v), "f"(x[i]));
        const float un = fma1(d, 0.5f, x[i]);
        const float s = fset(un);
        x[i] = fma1(-s, un, un);
      }
    } else if (MODE == 9) {
      const uint64_t rr = pack2((float)r, (float)r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t v = fma2(rr, a2, b2);
        uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(p[i]));
        const uint64_t un = fma2(d, h2, p[i]);
        float u0, u1; unpack2(un, u0, u1);
        const uint64_t s = pack2(-fset(u0), -fset(u1));
        p[i] = fma2(s, un, un);
      }
    } else if (MODE == 10) {
      const uint64_t rr = pack2((float)r, (float)r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        {
          const uint64_t v = fma2(rr, a2, b2);
          uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(p[i]));
          const uint64_t un = fma2(d, h2, p[i]);
          float u0, u1; unpack2(un, u0, u1);
          const uint64_t s = pack2(-fset(u0), -fset(u1));
          p[i] = fma2(s, un, un);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int n = 8 + 2 * i + q;
          const float v = fma1((float)r, a, b);
          float d; asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(v), "f"(x[n]));
          const float un = fma1(d, 0.5f, x[n]);
          const float s = fset(un);
          x[n] = fma1(-s, un, un);
        }
      }
    }
  }
  if (MODE >= 11 && MODE <= 13) {
    // 32 neurons per thread = 4 quad positions x 8 quads (4 packed pairs each), like k_conv1_umma<true,16>
    uint64_t u[4][4];
    float av[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) { u[j][q] = p[(j * 4 + q) & 7]; av[j][2 * q] = x[(j * 4 + q) & 15]; av[j][2 * q + 1] = x[(j * 3 + q) & 15]; }
    uint8_t *y = reinterpret_cast<uint8_t *>(out) + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4096 + (threadIdx.x & 31);
    constexpr size_t ystride = 128;
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t v = fma2(pack2(av[j][2 * q], av[j][2 * q + 1]), a2, b2);
          uint64_t un;
          if (MODE == 11 || MODE == 13) {
            uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(u[j][q]));
            un = fma2(d, h2, u[j][q]);
          } else {
            un = fma2(u[j][q], h2, v);
          }
          float u0, u1; unpack2(un, u0, u1);
          const float s0 = fset(u0), s1 = fset(u1);
          u[j][q] = fma2(pack2(-s0, -s1), un, un);
          w0[j] = __float_as_uint(s0); w1[j] = __float_as_uint(s1);
        }
        if (MODE != 13) {
          y[(2 * q) * ystride] = (uint8_t)(((w0[0] | w0[1]) | (w0[2] | w0[3])) >> 29);
          y[(2 * q + 1) * ystride] = (uint8_t)(((w1[0] | w1[1]) | (w1[2] | w1[3])) >> 29);
        } else {
          x[q] += __uint_as_float((w0[0] | w0[1]) | (w0[2] | w0[3])) + __uint_as_float((w1[0] | w1[1]) | (w1[2] | w1[3]));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) p[(j + q) & 7] ^= u[j][q];
  }
  if (MODE >= 14 && MODE <= 17) {
    // scaled-domain single-rounding LIF: W_t = fma(W_{t-1}, keep_{t-1}, V_t), keep_t = (W_t < 2^(t+1))
    // MODE 14: keep by FSET, pool by LOP3 + SHF | 15: keep by scalar FFMA.SAT, pool by LOP3 + SHF
    // MODE 16: keep by FSET, pool on the FMA pipe (product, low-byte trick) | 17: FFMA.SAT + FMA-pipe pool
    constexpr bool SAT = (MODE == 15 || MODE == 17), FPOOL = (MODE == 16 || MODE == 17);
    uint64_t w[4][4], kp[4][4];
    float av[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) { w[j][q] = p[(j * 4 + q) & 7]; kp[j][q] = pack2(1.f, 1.f); av[j][2 * q] = x[(j * 4 + q) & 15]; av[j][2 * q + 1] = x[(j * 3 + q) & 15]; }
    uint8_t *y = reinterpret_cast<uint8_t *>(out) + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4096 + (threadIdx.x & 31);
    float sct = a, bit = b, tht = 2.0f, big = 1e20f;
    const uint64_t eps2 = pack2(-1.1920929e-7f, -1.1920929e-7f), one2 = pack2(1.0f + 1.1920929e-7f, 1.0f + 1.1920929e-7f);
    for (int r = 0; r < R; ++r) {
      const uint64_t sc2 = pack2(sct, sct), bi2 = pack2(bit, bit);
      const float nbig = -big, bth = big * tht;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint64_t kk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t v = fma2(pack2(av[j][2 * q], av[j][2 * q + 1]), sc2, bi2);
          const uint64_t wn = fma2(w[j][q], kp[j][q], v);
          float u0, u1; unpack2(wn, u0, u1);
          float s0, s1;
          if (SAT) {
            asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(s0) : "f"(u0), "f"(nbig), "f"(bth));
            asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(s1) : "f"(u1), "f"(nbig), "f"(bth));
          } else {
            asm volatile("set.lt.f32.f32 %0, %1, %2;" : "=f"(s0) : "f"(u0), "f"(tht));
            asm volatile("set.lt.f32.f32 %0, %1, %2;" : "=f"(s1) : "f"(u1), "f"(tht));
          }
          w[j][q] = wn; kk[j] = kp[j][q] = pack2(s0, s1);
        }
        if (FPOOL) {
          uint64_t pr;
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(kk[0]), "l"(kk[1]));
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pr), "l"(kk[2]));
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pr), "l"(kk[3]));
          uint64_t z;
          z = fma2(pr, eps2, one2);
          float z0, z1; unpack2(z, z0, z1);
          y[(2 * q) * 128] = (uint8_t)__float_as_uint(z0);
          y[(2 * q + 1) * 128] = (uint8_t)__float_as_uint(z1);
        } else {
          float a0[4], a1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack2(kk[j], a0[j], a1[j]);
          y[(2 * q) * 128] = (uint8_t)((((__float_as_uint(a0[0]) & __float_as_uint(a0[1]) & __float_as_uint(a0[2])) & __float_as_uint(a0[3])) ^ 0x3F800000u) >> 29);
          y[(2 * q + 1) * 128] = (uint8_t)((((__float_as_uint(a1[0]) & __float_as_uint(a1[1]) & __float_as_uint(a1[2])) & __float_as_uint(a1[3])) ^ 0x3F800000u) >> 29);
        }
      }
      sct *= 1.0001f; bit *= 1.0001f; tht *= 1.0001f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) p[(j + q) & 7] ^= w[j][q] ^ kp[j][q];
  }
  if (MODE == 18) {
    uint64_t wb[4][4];
    float av[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) { wb[j][q] = p[(j * 4 + q) & 7]; av[j][2 * q] = x[(j * 4 + q) & 15]; av[j][2 * q + 1] = x[(j * 3 + q) & 15]; }
    uint8_t *y = reinterpret_cast<uint8_t *>(out) + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4096 + (threadIdx.x & 31);
    float sct = a, bit = b, big = 1e20f, tht = 2.0f;
    const uint64_t eps2 = pack2(-1.1920929e-7f, -1.1920929e-7f), one2 = pack2(1.0f + 1.1920929e-7f, 1.0f + 1.1920929e-7f);
    for (int r = 0; r < R; ++r) {
      const uint64_t sc2 = pack2(sct, sct);
      bit *= 1.0001f;
      const uint64_t bn2 = pack2(bit, bit);
      const float nbig = -big, bth = big * tht;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint64_t kk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t wn = fma2(pack2(av[j][2 * q], av[j][2 * q + 1]), sc2, wb[j][q]);
          float u0, u1; unpack2(wn, u0, u1);
          float s0, s1;
          asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(s0) : "f"(u0), "f"(nbig), "f"(bth));
          asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(s1) : "f"(u1), "f"(nbig), "f"(bth));
          kk[j] = pack2(s0, s1);
          wb[j][q] = fma2(wn, kk[j], bn2);
        }
        uint64_t pr;
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(kk[0]), "l"(kk[1]));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pr), "l"(kk[2]));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pr), "l"(kk[3]));
        const uint64_t z = fma2(pr, eps2, one2);
        float z0, z1; unpack2(z, z0, z1);
        y[(2 * q) * 128] = (uint8_t)__float_as_uint(z0);
        y[(2 * q + 1) * 128] = (uint8_t)__float_as_uint(z1);
      }
      sct *= 1.0001f; tht *= 1.0001f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) p[(j + q) & 7] ^= wb[j][q];
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float u0, u1; unpack2(p[i], u0, u1); acc += u0 + u1; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, int slots, float *out, long long *cyc) {
  for (int warps_per_smsp : {1, 2, 4}) {
    k<MODE><<<148, 128 * warps_per_smsp>>>(out, cyc, 1e-3f);
    k<MODE><<<148, 128 * warps_per_smsp>>>(out, cyc, 1e-3f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-44s warps/SMSP %d: %7.2f cycles per round per SMSP  (%d neuron-or-op slots/warp-round -> %.3f cyc/slot)\n", name,
           warps_per_smsp, (double)c / R, slots, (double)c / R / (slots * warps_per_smsp));
  }
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, (size_t)148 * 512 * 2048 + (1 << 20)); cudaMalloc(&cyc, 8);
  run<0>("16 scalar FFMA", 16, out, cyc);
  run<1>("8 FFMA2 (=16 lane-ops)", 16, out, cyc);
  run<2>("8 scalar FFMA + 4 FFMA2 (=16 lane-ops)", 16, out, cyc);
  run<3>("16 FSET.BF", 16, out, cyc);
  run<4>("16 FSETP+FSEL", 16, out, cyc);
  run<5>("16 FFMA + 8 FSET", 16, out, cyc);
  run<6>("8 FFMA2 + 8 FSET", 16, out, cyc);
  run<7>("8 FFMA + 4 FFMA2 + 8 FSET", 16, out, cyc);
  run<8>("LIF exact scalar x16 neurons", 16, out, cyc);
  run<9>("LIF exact packed x16 neurons", 16, out, cyc);
  run<10>("LIF exact half packed/half scalar x16", 16, out, cyc);
  run<11>("conv1 epilogue step (32 neurons, exact, OR-pool, STG.U8)", 32, out, cyc);
  run<12>("conv1 epilogue step (32 neurons, single-rounding form)", 32, out, cyc);
  run<13>("conv1 epilogue step (32 neurons, exact, no stores)", 32, out, cyc);
  run<14>("scaled 3-op LIF: FSET keep, LOP3 pool", 32, out, cyc);
  run<15>("scaled 3-op LIF: FFMA.SAT keep, LOP3 pool", 32, out, cyc);
  run<16>("scaled 3-op LIF: FSET keep, FMA-pipe pool", 32, out, cyc);
  run<17>("scaled 3-op LIF: FFMA.SAT keep, FMA-pipe pool", 32, out, cyc);
  run<18>("scaled 3-op LIF, one state register (bias folded)", 32, out, cyc);
  return 0;
}
