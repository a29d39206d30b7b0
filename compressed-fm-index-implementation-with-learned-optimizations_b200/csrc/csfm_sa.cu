// csfm_sa.cu — suffix array + BWT + sampled SA on the device (prefix doubling, radix sort).
//
// Replaces cs::build_sa_naive (/root/reference/src/core/sais.hpp:8-16, an O(n^2 log n) std::sort
// on substr copies), cs::build_bwt_from_sa (src/core/bwt.hpp:7-15) and the SSA loop of
// FMIndex::build_from_text (src/api/fm_index.cpp:57-65). The order is the reference's:
// unsigned-byte lexicographic, and a suffix that is a proper prefix of another sorts first
// (std::string::operator<). Suffixes are pairwise distinct, so that total order has exactly one
// suffix array — the output is bit-identical to the reference's wherever the reference can run.
//
// Round 0 packs the first k symbols of every suffix into a 64-bit key: symbols are recoded to
// 1..sigma (0 = past the end, which makes "shorter first" fall out of integer order), `bps` =
// ceil(log2(sigma+1)) bits each, k = 64/bps (21 symbols for DNA+$, 7 for a full byte alphabet).
// Every following round sorts by (rank[i], rank[i+h]+1 or 0 past the end) and doubles h, until
// all ranks are distinct. Sorting is cub::DeviceRadixSort on (u64 key, u32 suffix) pairs.
#include <algorithm>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "csfm_host.hpp"

namespace csfm {

namespace {

struct SymTable {
  uint16_t code[256];  // 1..sigma for present bytes (sigma may be 256), 0 = past the end
};

__global__ void byte_hist_kernel(const uint8_t* __restrict__ t, uint64_t n, unsigned int* __restrict__ present) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) sh[t[i]] = 1;
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) present[i] = 1;
}

__global__ void init_keys_kernel(const uint8_t* __restrict__ t, uint64_t n, int bps, int k,
                                 const __grid_constant__ SymTable st, uint64_t* __restrict__ key,
                                 uint32_t* __restrict__ sa) {
  __shared__ uint16_t sc[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sc[i] = st.code[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t kk = 0;
    for (int j = 0; j < k; ++j) {
      const uint64_t code = (i + j < n) ? sc[t[i + j]] : 0;
      kk = (kk << bps) | code;
    }
    key[i] = kk;
    sa[i] = (uint32_t)i;
  }
}

// heads[j] = j if key[j] starts a new group else 0; counts the groups.
__global__ void mark_heads_kernel(const uint64_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ heads,
                                  unsigned long long* __restrict__ ngroups) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned int local = 0;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const bool head = (j == 0) || (key[j] != key[j - 1]);
    heads[j] = head ? (uint32_t)j : 0u;
    local += head;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(ngroups, (unsigned long long)local);
}

__global__ void scatter_rank_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ grp, uint64_t n,
                                    uint32_t* __restrict__ rank) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) rank[sa[j]] = grp[j];
}

__global__ void doubling_keys_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint64_t n,
                                     uint64_t h, uint64_t* __restrict__ key) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint64_t s = sa[j];
    const uint64_t r2 = (s + h < n) ? (uint64_t)rank[s + h] + 1 : 0;  // shorter suffix first
    key[j] = ((uint64_t)rank[s] << 32) | r2;
  }
}

// BWT[j] = T[(SA[j]-1) mod n]  (bwt.hpp:10-13);  ssa[k] = SA[k*stride]  (fm_index.cpp:60-64)
__global__ void bwt_ssa_kernel(const uint8_t* __restrict__ t, const uint32_t* __restrict__ sa, uint64_t n,
                               uint32_t stride_s, uint8_t* __restrict__ bwt, uint32_t* __restrict__ ssa) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint32_t s = sa[j];
    bwt[j] = (s == 0) ? t[n - 1] : t[s - 1];
    if (j % stride_s == 0) ssa[j / stride_s] = s;
  }
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

}  // namespace

int build_sa_bwt_device(const uint8_t* d_text, uint64_t n, uint32_t stride, cudaStream_t stream,
                        uint8_t** d_bwt_out, uint32_t** d_ssa_out, uint64_t* nsamp_out,
                        uint32_t** d_sa_out, uint32_t* rounds_out, uint32_t* passes_out) {
  if (rounds_out) *rounds_out = 0;
  if (passes_out) *passes_out = 0;
  *d_bwt_out = nullptr;
  *d_ssa_out = nullptr;
  if (d_sa_out) *d_sa_out = nullptr;
  if (n > kMaxN) return fail(CSFM_ERR_TOO_LARGE, "text length must be < 2^32 - 1");
  if (stride == 0) return fail(CSFM_ERR_INVALID, "ssa_stride must be > 0");
  const uint64_t nsamp = (n + stride - 1) / stride;
  *nsamp_out = nsamp;
  if (n == 0) return CSFM_OK;

  uint64_t *key_a = nullptr, *key_b = nullptr;
  uint32_t *sa_a = nullptr, *sa_b = nullptr, *rank = nullptr, *heads = nullptr;
  unsigned int* d_present = nullptr;
  unsigned long long* d_ngroups = nullptr;
  void* d_tmp = nullptr;
  uint8_t* d_bwt = nullptr;
  uint32_t* d_ssa = nullptr;
  auto cleanup = [&]() {
    cudaFree(key_a); cudaFree(key_b); cudaFree(sa_a); cudaFree(sa_b); cudaFree(rank); cudaFree(heads);
    cudaFree(d_present); cudaFree(d_ngroups); cudaFree(d_tmp);
  };
#define SA_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cleanup();                                                                             \
      cudaFree(d_bwt);                                                                       \
      cudaFree(d_ssa);                                                                       \
      return fail(_e == cudaErrorMemoryAllocation ? CSFM_ERR_NOMEM : CSFM_ERR_CUDA,          \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
    }                                                                                        \
  } while (0)

  SA_CUDA(cudaMalloc(&key_a, n * 8));
  SA_CUDA(cudaMalloc(&key_b, n * 8));
  SA_CUDA(cudaMalloc(&sa_a, n * 4));
  SA_CUDA(cudaMalloc(&sa_b, n * 4));
  SA_CUDA(cudaMalloc(&rank, n * 4));
  SA_CUDA(cudaMalloc(&heads, n * 4));
  SA_CUDA(cudaMalloc(&d_present, 256 * 4));
  SA_CUDA(cudaMalloc(&d_ngroups, 8));

  // alphabet
  SA_CUDA(cudaMemsetAsync(d_present, 0, 256 * 4, stream));
  byte_hist_kernel<<<1024, 256, 0, stream>>>(d_text, n, d_present);
  unsigned int present[256];
  SA_CUDA(cudaMemcpyAsync(present, d_present, sizeof present, cudaMemcpyDeviceToHost, stream));
  SA_CUDA(cudaStreamSynchronize(stream));
  SymTable st;
  std::memset(st.code, 0, sizeof st.code);
  uint32_t sigma = 0;
  for (int c = 0; c < 256; ++c)
    if (present[c]) st.code[c] = (uint16_t)(++sigma);
  int bps = 1;
  while ((1u << bps) < sigma + 1) ++bps;  // codes 0..sigma
  const int k0 = 64 / bps;                // 21 symbols for DNA+$, 7 for a full byte alphabet

  const int grid = 148 * 8, block = 256;
  init_keys_kernel<<<grid, block, 0, stream>>>(d_text, n, bps, k0, st, key_a, sa_a);

  cub::DoubleBuffer<uint64_t> keys(key_a, key_b);
  cub::DoubleBuffer<uint32_t> vals(sa_a, sa_b);
  size_t tmp_sort = 0, tmp_scan = 0;
  SA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, vals, (int64_t)n, 0, 64, stream));
  SA_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, heads, heads, MaxOp(), (int64_t)n, stream));
  const size_t tmp_bytes = std::max(tmp_sort, tmp_scan) + 256;
  SA_CUDA(cudaMalloc(&d_tmp, tmp_bytes));

  int nbits_n = 1;
  while ((1ull << nbits_n) < n + 1) ++nbits_n;
  uint64_t h = (uint64_t)k0;
  int end_bit = bps * k0;
  for (int round = 0;; ++round) {
    size_t tb = tmp_bytes;
    SA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys, vals, (int64_t)n, 0, end_bit, stream));
    if (rounds_out) ++*rounds_out;
    if (passes_out) *passes_out += (uint32_t)((end_bit + 7) / 8);  // 8-bit digits: each pass reads and writes every (key, suffix) pair
    SA_CUDA(cudaMemsetAsync(d_ngroups, 0, 8, stream));
    mark_heads_kernel<<<grid, block, 0, stream>>>(keys.Current(), n, heads, d_ngroups);
    unsigned long long ngroups = 0;
    SA_CUDA(cudaMemcpyAsync(&ngroups, d_ngroups, 8, cudaMemcpyDeviceToHost, stream));
    SA_CUDA(cudaStreamSynchronize(stream));
    if (ngroups == n) break;
    tb = tmp_bytes;
    SA_CUDA(cub::DeviceScan::InclusiveScan(d_tmp, tb, heads, heads, MaxOp(), (int64_t)n, stream));
    scatter_rank_kernel<<<grid, block, 0, stream>>>(vals.Current(), heads, n, rank);
    doubling_keys_kernel<<<grid, block, 0, stream>>>(vals.Current(), rank, n, h, keys.Current());
    h *= 2;
    end_bit = 32 + nbits_n;
    if (round > 40) {
      cleanup();
      return fail(CSFM_ERR_CUDA, "suffix sorting did not converge");
    }
  }
  const uint32_t* d_sa = vals.Current();
  SA_CUDA(cudaMalloc(&d_bwt, n));
  SA_CUDA(cudaMalloc(&d_ssa, nsamp * 4));
  bwt_ssa_kernel<<<grid, block, 0, stream>>>(d_text, d_sa, n, stride, d_bwt, d_ssa);
  SA_CUDA(cudaGetLastError());
  SA_CUDA(cudaStreamSynchronize(stream));
  if (d_sa_out) {  // hand the SA buffer over instead of freeing it
    if (d_sa == sa_a) { *d_sa_out = sa_a; sa_a = nullptr; } else { *d_sa_out = sa_b; sa_b = nullptr; }
  }
  cleanup();
#undef SA_CUDA
  *d_bwt_out = d_bwt;
  *d_ssa_out = d_ssa;
  return CSFM_OK;
}

}  // namespace csfm
