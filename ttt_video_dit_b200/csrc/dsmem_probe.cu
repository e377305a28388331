// Distributed-shared-memory microbenchmark (self-test library only): one-way time of moving `bytes` from one CTA of a
// 2-CTA cluster to the other, including the completion signal, measured as a ping-pong (cycles / 2).
//   mode 0: cp.async.bulk.shared::cluster.shared::cta (TMA engine, completes tx bytes on the peer's mbarrier)
//   mode 1: st.shared::cluster.v4 by 256 threads + one remote mbarrier arrive per thread (release.cluster)
//   mode 2: as mode 0 but the payload is split into 4 bulk copies issued back to back
// Used to decide whether a hidden-dimension split of the TTT scans across a CTA pair can afford its per-step exchanges.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "ttt_internal.h"

namespace tb {

__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
dsmem_probe_kernel(int mode, int bytes, int iters, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1;
  const int tid = threadIdx.x;
  const uint32_t buf = smem_u32(smem);              // [0, bytes): receive buffer ; [bytes, 2*bytes): send buffer
  const uint32_t peer_buf = mapa(buf, peer), peer_bar = mapa(smem_u32(&bar), peer);
  if (tid == 0) {
    mbar_init(&bar, mode == 1 ? 256 : 1);
    fence_mbar_init();
  }
  for (int i = tid; i < bytes / 4; i += 256) reinterpret_cast<uint32_t*>(smem + bytes)[i] = i;
  fence_proxy_async();
  __syncthreads();
  cluster_sync_all();
  uint32_t phase = 0;
  long long t0 = 0;
  auto send = [&]() {
    if (mode == 1) {
      for (int c = tid; c < bytes / 16; c += 256) {
        uint32_t a, b, cc, d;
        ld_shared_v4(buf + bytes + c * 16, a, b, cc, d);
        asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(peer_buf + c * 16), "r"(a), "r"(b), "r"(cc), "r"(d) : "memory");
      }
      remote_arrive(peer_bar);
    } else if (tid == 0) {
      const int parts = mode == 2 ? 4 : 1;
      for (int q = 0; q < parts; ++q)
        bulk_copy_to_peer(peer_buf + q * (bytes / parts), buf + bytes + q * (bytes / parts), bytes / parts, peer_bar);
    }
  };
  auto arm = [&]() {
    if (mode != 1 && tid == 0) mbar_expect_tx(&bar, (uint32_t)bytes);
  };
  if (rank == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      arm();
      send();
      wait_cluster(&bar, phase);
      phase ^= 1;
      __syncthreads();
    }
    if (tid == 0) out[0] = (float)(clock64() - t0) / (2.0f * iters);
  } else {
    for (int i = 0; i < iters; ++i) {
      arm();
      wait_cluster(&bar, phase);
      phase ^= 1;
      __syncthreads();
      send();
    }
  }
  cluster_sync_all();
}

cudaError_t launch_dsmem_probe(int mode, int bytes, int iters, float* out, cudaStream_t stream) {
  if (mode < 0 || mode > 2 || bytes < 64 || bytes > 98304 || bytes % 64 || iters <= 0) return cudaErrorInvalidValue;
  const size_t smem = 2 * (size_t)bytes + 1024;
  TB_TRY(cudaFuncSetAttribute(dsmem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr (dsmem probe)");
  dsmem_probe_kernel<<<2, 256, smem, stream>>>(mode, bytes, iters, out);
  return cudaGetLastError();
}

}  // namespace tb
