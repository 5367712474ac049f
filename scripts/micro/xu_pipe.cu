// Micro-benchmark: which pipe do the softmax pass's instructions occupy?  One warp per SM sub-partition (128 threads per CTA,
// one CTA per SM) runs straight-line sequences and times them with clock64:
//   A  128 MUFU.EX2                          B  64 F2FP (cvt.rn.f16x2.f32)          C  128 MUFU.EX2 + 64 F2FP interleaved
//   D  the pass as the kernel issues it: per pair FFMA2, 2 MUFU.EX2, FADD2, F2FP
//   E  64 pairs of a degree-4 exp2 polynomial on the FMA pipe (Cody-Waite, packed fp32x2), no MUFU
//   F  pass with every third pair's exponentials on the FMA pipe
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/xu_pipe scripts/micro/xu_pipe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float hi, float lo) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// 2^t for a pair, t <= 0: n = round(t), f = t - n in [-0.5, 0.5], 2^f by a degree-4 polynomial, exponent added as an integer
__device__ __forceinline__ void poly_ex2_pair(uint64_t t2, float& p0, float& p1) {
  const uint64_t magic = pk(12582912.f, 12582912.f), nmagic = pk(-12582912.f, -12582912.f), none = pk(-1.f, -1.f);
  const uint64_t lo = pk(-125.f, -125.f);
  float a, b;
  upk(t2, a, b);
  t2 = pk(fmaxf(a, -125.f), fmaxf(b, -125.f));
  const uint64_t r = add2(t2, magic);                    // integer part in the low mantissa bits
  const uint64_t n = add2(r, nmagic);
  const uint64_t f = fma2(n, none, t2);                  // t - n
  uint64_t y = pk(9.6181291e-3f, 9.6181291e-3f);
  y = fma2(y, f, pk(5.5504109e-2f, 5.5504109e-2f));
  y = fma2(y, f, pk(2.4022651e-1f, 2.4022651e-1f));
  y = fma2(y, f, pk(6.9314718e-1f, 6.9314718e-1f));
  y = fma2(y, f, pk(1.f, 1.f));
  float y0, y1, r0, r1;
  upk(y, y0, y1);
  upk(r, r0, r1);
  p0 = __uint_as_float(__float_as_uint(y0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(y1) + (__float_as_uint(r1) << 23));
  (void)lo;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(const float* in, float* out, long long* cyc) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[i * 128 + threadIdx.x];
  uint32_t ph[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) ph[i] = 0;
  uint64_t sum2 = pk(0.f, 0.f);
  const uint64_t sc = pk(0.147f, 0.147f), nm = pk(-1.25f, -1.25f);
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = ex2(s[i]);
  } else if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 64; ++i) ph[i] = cvt2(s[2 * i + 1], s[2 * i]);
  } else if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < 64; ++i) { s[2 * i] = ex2(s[2 * i]); s[2 * i + 1] = ex2(s[2 * i + 1]); ph[i] = cvt2(s[(2 * i + 65) & 127], s[(2 * i + 64) & 127]); }
  } else if (MODE == 3 || MODE == 5) {
    float pb[2][8];
    auto grp = [&](int q, float* dst) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = 8 * q + 2 * i;
        const uint64_t t2 = fma2(pk(s[kk], s[kk + 1]), sc, nm);
        if (MODE == 5 && ((4 * q + i) % 3 == 2)) {
          poly_ex2_pair(t2, dst[2 * i], dst[2 * i + 1]);
        } else {
          float t0f, t1f;
          upk(t2, t0f, t1f);
          dst[2 * i] = ex2(t0f);
          dst[2 * i + 1] = ex2(t1f);
        }
      }
    };
    grp(0, pb[0]);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (q + 1 < 16) grp(q + 1, pb[(q + 1) & 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sum2 = add2(sum2, pk(pb[q & 1][2 * i], pb[q & 1][2 * i + 1]));
        ph[4 * q + i] = cvt2(pb[q & 1][2 * i + 1], pb[q & 1][2 * i]);
      }
    }
  } else if (MODE == 4) {
#pragma unroll
    for (int i = 0; i < 64; ++i) poly_ex2_pair(fma2(pk(s[2 * i], s[2 * i + 1]), sc, nm), s[2 * i], s[2 * i + 1]);
  }
  const long long t1 = clock64();
  float acc = 0.f, a, b;
  upk(sum2, a, b);
#pragma unroll
  for (int i = 0; i < 128; ++i) acc += s[i];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc += __uint_as_float(ph[i]);
  out[blockIdx.x * 128 + threadIdx.x] = acc + a + b;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 65536); cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&cyc, 64);
  static float h[16384];
  for (int i = 0; i < 16384; ++i) h[i] = -0.001f * (i % 7919);
  cudaMemcpy(in, h, 65536, cudaMemcpyHostToDevice);
  const char* names[6] = {"A 128 MUFU.EX2", "B 64 F2FP", "C 128 MUFU.EX2 + 64 F2FP", "D pass (FFMA2, 2 MUFU, FADD2, F2FP per pair)",
                          "E 64 pairs polynomial exp2 (FMA pipe)", "F pass, every third pair polynomial"};
  for (int rep = 0; rep < 2; ++rep)
    for (int m = 0; m < 6; ++m) {
      switch (m) {
        case 0: k<0><<<148, 128>>>(in, out, cyc); break;
        case 1: k<1><<<148, 128>>>(in, out, cyc); break;
        case 2: k<2><<<148, 128>>>(in, out, cyc); break;
        case 3: k<3><<<148, 128>>>(in, out, cyc); break;
        case 4: k<4><<<148, 128>>>(in, out, cyc); break;
        case 5: k<5><<<148, 128>>>(in, out, cyc); break;
      }
      long long c[4];
      cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
      if (rep == 1) printf("%-52s %lld cycles (one warp per sub-partition; warps: %lld %lld %lld %lld)\n", names[m], c[0], c[0], c[1], c[2], c[3]);
    }
  // accuracy of the polynomial against exp2f
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
