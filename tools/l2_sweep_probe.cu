// tools/l2_sweep_probe.cu — random-fetch rate of the memory system as a function of the working-set size
// (16 MiB .. 1 GiB) and of the unit fetched (32 / 64 / 128 bytes), dependent chains like the query kernels
// issue them (one unit per sub-warp of unit/32 lanes, 256-bit loads). Tells which index sizes are served
// at L2 rate on B200 (two L2 partitions: a line may be cached in both) and what the L2-regime ceiling is
// for the roofline of the L2-resident configurations (C1, C2, C4). A measurement tool, not product code.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int LANES>  // LANES x 32 B per unit
__global__ void chase_kernel(const uint8_t* __restrict__ buf, uint32_t nunits, int iters, uint32_t* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int j = lane % LANES;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  uint64_t s = group * 0x9E3779B97F4A7C15ull + 0x1234567;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    const uint32_t u = (uint32_t)((((s >> 32) + acc) * (uint64_t)nunits) >> 32);
    uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7)
                 : "l"(buf + ((uint64_t)u * LANES + j) * 32));
    uint32_t t = v0 ^ v1 ^ v2 ^ v3 ^ v4 ^ v5 ^ v6 ^ v7;
    for (int o = 1; o < LANES; o <<= 1) t ^= __shfl_xor_sync(0xFFFFFFFFu, t, o);
    acc += t & 1;  // buffer is zero: acc stays 0 but the dependency is real
  }
  if (acc == 0x12345) sink[0] = acc;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"results\": [\n", p.name, p.multiProcessorCount, p.l2CacheSize);
  const size_t big = 1ull << 30;
  uint8_t* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 0, big));
  uint32_t* sink; CK(cudaMalloc(&sink, 4));
  const int block = 256;
  bool first = true;
  for (int ctas : {8, 4}) {
    const int grid = p.multiProcessorCount * ctas;
    for (size_t mib : {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 384, 512, 1024}) {
      const size_t bytes = mib << 20;
      const int iters = 256;
      const double threads = (double)grid * block;
      auto run = [&](int unit, float ms, double loads) {
        printf("%s  {\"ctas_per_sm\": %d, \"buffer_mib\": %zu, \"unit_bytes\": %d, \"ms\": %.3f, \"gunits_per_s\": %.2f, \"gbytes_per_s\": %.1f}",
               first ? "" : ",\n", ctas, mib, unit, ms, loads / ms / 1e6, loads * unit / ms / 1e6);
        first = false;
      };
      { float ms = time_ms([&] { chase_kernel<1><<<grid, block>>>(buf, (uint32_t)(bytes / 32), iters, sink); }); run(32, ms, threads * iters); }
      { float ms = time_ms([&] { chase_kernel<2><<<grid, block>>>(buf, (uint32_t)(bytes / 64), iters, sink); }); run(64, ms, threads / 2 * iters); }
      { float ms = time_ms([&] { chase_kernel<4><<<grid, block>>>(buf, (uint32_t)(bytes / 128), iters, sink); }); run(128, ms, threads / 4 * iters); }
    }
  }
  printf("\n]}\n");
  return 0;
}
