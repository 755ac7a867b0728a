// stream_probe.cu -- DEV TOOL (not product): how fast can a B200 SM pull a dense byte stream into shared
// memory / registers?  Variants of the staging scheme used by k_stream_tma, with a trivial consumer, to
// separate the memory side of that kernel from its ALU side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu && ./stream_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TMA ring: tileBytes per stage, `split` bulk copies per tile, `stages` deep; consumers XOR a few words
template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1) k_tma(const uint8_t *__restrict__ src, size_t bytes, uint32_t tileBytes,
                                                   int stages, int split, uint32_t *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm), *empty = full + 8;
  unsigned char *buf = sm + 128;
  const uint32_t nTiles = (uint32_t) (bytes / tileBytes), tid = threadIdx.x;
  auto issue = [&](uint32_t stage, uint32_t tile) {
    mbar_expect_tx(&full[stage], tileBytes);
    const uint32_t part = tileBytes / split;
    for (int s = 0; s < split; ++s) bulk_g2s(buf + (size_t) stage * tileBytes + s * part, src + (size_t) tile * tileBytes + s * part, part, &full[stage]);
  };
  if (tid == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kThreads / 32); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) for (int s = 0; s < stages; ++s) { uint32_t t = blockIdx.x + s * gridDim.x; if (t < nTiles) issue(s, t); }
  uint32_t acc = 0, k = 0;
  for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++k) {
    const uint32_t stage = k % stages, parity = (k / stages) & 1;
    if (tid == 0 && k > 0) {
      const uint32_t ps = (k - 1) % stages, pp = ((k - 1) / stages) & 1, nt = tile + (stages - 1) * gridDim.x;
      if (nt < nTiles) { mbar_wait(&empty[ps], pp); issue(ps, nt); }
    }
    mbar_wait(&full[stage], parity);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + (size_t) stage * tileBytes);
    for (uint32_t i = tid; i < tileBytes / 4; i += kThreads) acc ^= w[i];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// plain vectorised loads, grid-stride, `unroll` 16-byte loads in flight per thread
template <int kUnroll>
__global__ void __launch_bounds__(256) k_ldg(const uint4 *__restrict__ src, size_t n16, uint32_t *out) {
  uint32_t acc = 0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < n16; i += kUnroll * stride) {
    uint4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

int main() {
  const size_t bytes = (size_t) 540 << 20;
  uint8_t *d; uint32_t *o;
  cudaMalloc(&d, bytes + (1 << 20)); cudaMalloc(&o, 64);
  cudaMemset(d, 1, bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char *name, auto launch) {
    for (int w = 0; w < 3; ++w) launch();
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    printf("%-44s %8.3f ms  %8.1f GB/s  %s\n", name, ms / 10, bytes / (ms / 10 * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  cudaFuncSetAttribute(k_tma<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_tma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Cfg { uint32_t tile; int stages, split; } cfgs[] = {
      {32768, 4, 1}, {32768, 4, 4}, {32768, 4, 16}, {32768, 6, 1}, {16384, 8, 1}, {16384, 12, 1}, {8192, 16, 1}, {8192, 24, 1}, {65536, 3, 1}, {65536, 3, 8}};
  for (auto c : cfgs) {
    char name[96]; snprintf(name, sizeof(name), "tma 1024thr tile %u stages %d split %d", c.tile, c.stages, c.split);
    size_t smem = 128 + (size_t) c.tile * c.stages;
    time(name, [&]() { k_tma<1024><<<148, 1024, smem>>>(d, bytes, c.tile, c.stages, c.split, o); });
    snprintf(name, sizeof(name), "tma  256thr tile %u stages %d split %d", c.tile, c.stages, c.split);
    time(name, [&]() { k_tma<256><<<148, 256, smem>>>(d, bytes, c.tile, c.stages, c.split, o); });
  }
  time("ldg.128 x1, 148*8 blocks", [&]() { k_ldg<1><<<148 * 8, 256>>>((const uint4 *) d, bytes / 16, o); });
  time("ldg.128 x4, 148*8 blocks", [&]() { k_ldg<4><<<148 * 8, 256>>>((const uint4 *) d, bytes / 16, o); });
  time("ldg.128 x8, 148*8 blocks", [&]() { k_ldg<8><<<148 * 8, 256>>>((const uint4 *) d, bytes / 16, o); });
  time("ldg.128 x4, 148*16 blocks", [&]() { k_ldg<4><<<148 * 16, 256>>>((const uint4 *) d, bytes / 16, o); });
  return 0;
}
