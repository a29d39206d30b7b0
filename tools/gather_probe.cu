// tools/gather_probe.cu — random-access ceilings of the memory system:
// independent gathers of G-byte units (G = 32, 64, 128) from a buffer of B bytes, one unit per
// sub-warp of G/16 lanes (128-bit loads), addresses from a per-thread xorshift. Reports
// units/s and GB/s for an HBM-resident (4 GiB) and an L2-resident (32 MiB) buffer, and for each
// value of cudaLimitMaxL2FetchGranularity. Not part of the product; a measurement tool.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int LANES, int ILP>  // LANES x 16 B per unit, ILP independent units in flight per sub-warp
__global__ void gather_kernel(const uint4* __restrict__ buf, uint64_t nunits, int iters, uint32_t* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int j = lane % LANES;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  uint64_t s = group * 0x9E3779B97F4A7C15ull + 0x1234567;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint4 v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;  // LCG: one IMAD per address
      const uint64_t u = (s >> 24) & (nunits - 1);              // nunits is a power of two
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(buf + u * LANES + j));
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  }
  if (acc == 0x12345) sink[0] = acc;
}

// dependent chain: next address derived from the loaded data (like the level-to-level chain)
template <int LANES>
__global__ void chase_kernel(const uint4* __restrict__ buf, uint64_t nunits, int iters, uint32_t* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int j = lane % LANES;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  uint64_t s = group * 0x9E3779B97F4A7C15ull + 0x1234567;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    const uint64_t u = ((s >> 24) + acc) & (nunits - 1);
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + u * LANES + j));
    uint32_t t = v.x ^ v.y ^ v.z ^ v.w;
    for (int o = 1; o < LANES; o <<= 1) t ^= __shfl_xor_sync(0xFFFFFFFFu, t, o);
    acc += t & 1;  // buffer is zero: acc stays 0 but the dependency is real
  }
  if (acc == 0x12345) sink[0] = acc;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
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
  size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"default_l2_fetch_granularity\": %zu, \"results\": [\n", p.name,
         p.multiProcessorCount, p.l2CacheSize, gran);
  const size_t big = 4ull << 30, small = 32ull << 20;
  uint4* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 0, big));
  uint32_t* sink; CK(cudaMalloc(&sink, 4));
  const int grid = p.multiProcessorCount * 8, block = 256;
  bool first = true;
  for (size_t g : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
    if (g) { if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g) != cudaSuccess) { cudaGetLastError(); continue; } }
    size_t now = 0; cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
    for (size_t bytes : {big, small}) {
      auto run = [&](const char* name, int unit, double loads, float ms) {
        printf("%s  {\"l2_fetch_granularity\": %zu, \"buffer_mib\": %zu, \"kernel\": \"%s\", \"unit_bytes\": %d, \"ms\": %.3f, "
               "\"gunits_per_s\": %.3f, \"gbytes_per_s\": %.1f}", first ? "" : ",\n", now, bytes >> 20, name, unit, ms,
               loads / ms / 1e6, loads * unit / ms / 1e6);
        first = false;
      };
      const int iters = 64;
      const double threads = (double)grid * block;
      { float ms = time_ms([&] { gather_kernel<2, 4><<<grid, block>>>(buf, bytes / 32, iters, sink); }); run("gather_ilp4", 32, threads / 2 * iters * 4, ms); }
      { float ms = time_ms([&] { gather_kernel<4, 4><<<grid, block>>>(buf, bytes / 64, iters, sink); }); run("gather_ilp4", 64, threads / 4 * iters * 4, ms); }
      { float ms = time_ms([&] { gather_kernel<8, 4><<<grid, block>>>(buf, bytes / 128, iters, sink); }); run("gather_ilp4", 128, threads / 8 * iters * 4, ms); }
      { float ms = time_ms([&] { gather_kernel<1, 4><<<grid, block>>>(buf, bytes / 16, iters, sink); }); run("gather_ilp4", 16, threads * iters * 4, ms); }
      { float ms = time_ms([&] { chase_kernel<2><<<grid, block>>>(buf, bytes / 32, iters * 4, sink); }); run("chase", 32, threads / 2 * iters * 4, ms); }
      { float ms = time_ms([&] { chase_kernel<4><<<grid, block>>>(buf, bytes / 64, iters * 4, sink); }); run("chase", 64, threads / 4 * iters * 4, ms); }
      { float ms = time_ms([&] { chase_kernel<1><<<grid, block>>>(buf, bytes / 16, iters * 4, sink); }); run("chase", 16, threads * iters * 4, ms); }
    }
  }
  printf("\n]}\n");
  return 0;
}
