// Micro-benchmark: cycles per tcgen05.mma for the shapes the attention kernel issues, as a function of the dependency pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cacophony_b200/csrc -I include -o gpurun_out/mma_chain scripts/micro/mma_chain.cu
// One CTA per SM (148 CTAs, all doing the same thing, so shared resources see the real load); thread 32 issues `n` MMAs in the
// given pattern, commits to an mbarrier and waits; clock64 around issue and around completion.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace caco;

enum Mode { QK_SS = 0, PV_TS = 1, PV_SS = 2, PV_TS_KMAJOR = 3 };

struct Res { long long issue, total; };

template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) chain_kernel(int n, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  __shared__ __align__(8) uint64_t bar_storage[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar = smem_u32(&bar_storage[0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_slot), 512); tmem_relinquish<1>(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 1 && lane == 0) {
    // idesc: M = 128, N
    const uint32_t idesc_k = umma_idesc_f16(128, N);                 // both K-major
    const uint32_t idesc_mn = umma_idesc_f16(128, N, false, true);   // B MN-major
    const uint64_t a_desc = umma_desc_kmajor_sw128(sb);               // A: 128 rows x 64 cols (16 KB)
    const uint64_t b_desc = umma_desc_kmajor_sw128(sb + 16384);       // B K-major: N rows x 64 cols
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        const uint32_t d = tm + (i % NACC) * N;                       // accumulator column base (NACC independent chains)
        const int ks = i & 3;
        if (MODE == QK_SS) {
          umma_f16<1>(d, a_desc + 2 * ks, b_desc + 2 * ks, idesc_k, 1u);
        } else if (MODE == PV_TS) {
          const uint64_t vb = umma_desc_mnmajor_sw128(sb + 16384 + (i & 7) * 2048, 16384);
          umma_f16_ts<1>(d, tm + 448 + (i & 7) * 8, vb, idesc_mn, 1u);
        } else if (MODE == PV_SS) {
          const uint64_t vb = umma_desc_mnmajor_sw128(sb + 16384 + (i & 7) * 2048, 16384);
          umma_f16<1>(d, a_desc + 2 * ks, vb, idesc_mn, 1u);
        } else {
          umma_f16_ts<1>(d, tm + 448 + (i & 7) * 8, b_desc + 2 * ks, idesc_k, 1u);
        }
      }
      const long long t1 = clock64();
      umma_commit<1>(bar);
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      if (rep == 2 && blockIdx.x == 0) { out->issue = t1 - t0; out->total = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tm, 512);
}

template <int MODE, int N, int NACC>
void run(const char* name, Res* d_out) {
  auto k = chain_kernel<MODE, N, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  for (int n : {8, 16, 64, 256}) {
    k<<<148, 128, 66560>>>(n, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    Res r;
    cudaMemcpy(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d acc=%d n=%3d  issue %6lld  total %6lld  per-mma %.1f  (%s)\n", name, N, NACC, n, r.issue, r.total,
           (double)r.total / n, cudaGetErrorString(e));
  }
}

int main() {
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  run<QK_SS, 128, 1>("QK SS kmajor", d);
  run<QK_SS, 128, 2>("QK SS kmajor", d);
  run<QK_SS, 64, 1>("QK SS kmajor", d);
  run<QK_SS, 64, 2>("QK SS kmajor", d);
  run<QK_SS, 256, 1>("QK SS kmajor", d);
  run<PV_TS, 96, 1>("PV TS, V mn-major", d);
  run<PV_TS, 96, 2>("PV TS, V mn-major", d);
  run<PV_TS, 48, 2>("PV TS, V mn-major", d);
  run<PV_TS, 128, 1>("PV TS, V mn-major", d);
  run<PV_SS, 96, 1>("PV SS, V mn-major", d);
  run<PV_TS_KMAJOR, 96, 1>("PV TS, V k-major", d);
  run<PV_TS_KMAJOR, 96, 2>("PV TS, V k-major", d);
  run<PV_TS_KMAJOR, 128, 1>("PV TS, V k-major", d);
  return 0;
}
