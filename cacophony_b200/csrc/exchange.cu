// K7 + the path's one exchange step, fused: e / ||e + 1e-10|| (caco.py:146,173) written straight into EVERY rank's gathered
// embedding matrix over NVLink peer mappings, followed by a release flag — the all-gather of SURVEY.md 8e without a
// collective launch.  A rank's rows are produced by one small kernel at the end of a tower; the data movement is the
// kernel's own 128-bit stores to peer addresses (NVSwitch gives every peer full bandwidth; 0.79 MB per rank and modality),
// so nothing has to be scheduled next to the towers' persistent kernels, nobody spins on an SM while a peer is late, and the
// only wait is a 1-CTA kernel right before the similarity launch that reads the gathered rows.
//
// Protocol (per modality, double-buffered by step parity): rank r writes rows [r*B, (r+1)*B) of buffer (step & 1) on every
// peer, fences at system scope, and the LAST block to finish stores flag[r] = step + 1 on every peer (release.sys).
// caco_wait_flags spins (acquire.sys, with a time-out that reports instead of hanging) until all `world` flags of the local
// rank reached step + 1.  A rank can run at most one step ahead of a peer (its step s+1 wait needs the peer's step s+1
// flag, which the peer stores after its own step-s similarity in stream order), so two buffers suffice.
#include <cuda_runtime.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int XCH_MAX_WORLD = 16;

__global__ void __launch_bounds__(256)
l2norm_scatter_kernel(const float* __restrict__ in, int rows, int dim, float eps, void* const* __restrict__ peer_base,
                      long long dst_byte_offset, long long flag_byte_offset, int flag_index, unsigned int epoch, int world,
                      unsigned int* __restrict__ ticket) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row < rows) {
    const float4* x = reinterpret_cast<const float4*>(in + (size_t)row * dim);
    const int nv = dim / 4;
    // the norm is summed in exactly the order l2norm_kernel (rowops.cu) uses — lane-strided single floats — so that the
    // sharded path stays BIT-equal to the single-GPU one (scripts/check_sharded.py); the division below is elementwise
    const float* xs = in + (size_t)row * dim;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) { const float v = xs[i] + eps; s = fmaf(v, v, s); }   // caco.py:146: || e + 1e-10 ||
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    for (int i = lane; i < nv; i += 32) {
      const float4 v = x[i];
      const float4 y = make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
      for (int p = 0; p < world; ++p) {
        float4* dst = reinterpret_cast<float4*>(static_cast<char*>(peer_base[p]) + dst_byte_offset) + (size_t)row * nv + i;
        *dst = y;
      }
    }
  }
  // publish: every thread's peer stores are ordered before its block's ticket, the last block's flag stores after all tickets
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      __threadfence();
      *ticket = 0u;                                            // ready for the next launch on this stream
      for (int p = 0; p < world; ++p) {
        unsigned int* f = reinterpret_cast<unsigned int*>(static_cast<char*>(peer_base[p]) + flag_byte_offset) + flag_index;
        st_release_sys_u32(f, epoch);
      }
    }
  }
}

__global__ void wait_flags_kernel(const unsigned int* __restrict__ flags, int n, unsigned int epoch, unsigned long long timeout_ns,
                                  int* __restrict__ status) {
  const int i = threadIdx.x;
  if (i >= n) return;
  const unsigned long long t0 = global_timer_ns();
  while ((int)(ld_acquire_sys_u32(flags + i) - epoch) < 0) {
    if (global_timer_ns() - t0 > timeout_ns) {                 // a peer never arrived: report, do not hang the GPU
      if (status) atomicExch(status, 1 + i);
      return;
    }
    __nanosleep(200);
  }
}

}  // namespace caco

extern "C" int caco_l2norm_scatter(const float* in, int rows, int dim, float eps, void* const* peer_base_dev,
                                   long long dst_byte_offset, long long flag_byte_offset, int flag_index, unsigned int epoch,
                                   int world, unsigned int* ticket, void* stream) {
  using namespace caco;
  if (!in || !peer_base_dev || !ticket || rows <= 0 || dim <= 0 || (dim & 3) || world < 1 || world > XCH_MAX_WORLD)
    return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (dst_byte_offset & 15) || (flag_byte_offset & 3)) return CACO_ERR_ALIGN;
  l2norm_scatter_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(in, rows, dim, eps, peer_base_dev, dst_byte_offset,
                                                                         flag_byte_offset, flag_index, epoch, world, ticket);
  count_launch();
  return (int)cudaGetLastError();
}

extern "C" int caco_wait_flags(const unsigned int* flags, int n, unsigned int epoch, int timeout_ms, int* status, void* stream) {
  using namespace caco;
  if (!flags || n < 1 || n > XCH_MAX_WORLD) return CACO_ERR_ARG;
  wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, epoch, (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ull,
                                                        status);
  count_launch();
  return (int)cudaGetLastError();
}
