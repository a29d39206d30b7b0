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
// Every following round doubles h and sorts by (rank[i], rank[i+h]+1 or 0 past the end) — but only the
// suffixes whose group still has more than one member: a suffix alone in its group is in its final row and
// is left where it is, the others are compacted (in row order), sorted, and written back into the rows they
// came from (the group rank in the high key bits keeps every group in its own run of rows). Random text is
// done after round 0 or 1; a text with long repeats keeps re-sorting only the suffixes inside the repeats,
// log2(repeat / k) times. Sorting is cub::DeviceRadixSort on (u64 key, u32 suffix) pairs.
#include <algorithm>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

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

// flag[j] = 1 if row j belongs to a group of more than one suffix (heads[] as written by mark_heads_kernel,
// BEFORE the scan: heads[j] == j for a group head, 0 otherwise; row 0 is always a head)
__global__ void unresolved_flags_kernel(const uint64_t* __restrict__ key, uint64_t n, uint8_t* __restrict__ flag) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const bool head = (j == 0) || (key[j] != key[j - 1]);
    const bool next_head = (j + 1 == n) || (key[j + 1] != key[j]);
    flag[j] = (head && next_head) ? 0 : 1;
  }
}

// keys of the unresolved rows for the next round: (rank[s], rank[s + h] + 1, or 0 past the end), s = sa[pos[i]]
__global__ void doubling_keys_sparse_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ pos,
                                            const uint32_t* __restrict__ rank, uint64_t n, uint64_t u, uint64_t h,
                                            uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < u; i += stride) {
    const uint64_t s = sa[pos[i]];
    const uint64_t r2 = (s + h < n) ? (uint64_t)rank[s + h] + 1 : 0;  // shorter suffix first
    key[i] = ((uint64_t)rank[s] << 32) | r2;
    val[i] = (uint32_t)s;
  }
}

// after the sort: run heads among the unresolved (index into the compacted order), and which of them stay unresolved
__global__ void sparse_heads_kernel(const uint64_t* __restrict__ key, uint64_t u, uint32_t* __restrict__ heads,
                                    uint8_t* __restrict__ flag) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < u; i += stride) {
    const bool head = (i == 0) || (key[i] != key[i - 1]);
    const bool next_head = (i + 1 == u) || (key[i + 1] != key[i]);
    heads[i] = head ? (uint32_t)i : 0u;
    flag[i] = (head && next_head) ? 0 : 1;
  }
}

// write the sorted suffixes back into their rows and give them the row of their group's head as rank
__global__ void sparse_scatter_kernel(const uint32_t* __restrict__ val, const uint32_t* __restrict__ pos,
                                      const uint32_t* __restrict__ head_of, uint64_t u, uint32_t* __restrict__ sa,
                                      uint32_t* __restrict__ rank) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < u; i += stride) {
    const uint32_t s = val[i];
    sa[pos[i]] = s;
    rank[s] = pos[head_of[i]];
  }
}

__global__ void iota_kernel(uint32_t* __restrict__ out, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// pos_out[scan[j]] = pos_in ? pos_in[j] : j for every flagged j (scan = exclusive prefix sum of the flags)
__global__ void compact_rows_kernel(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ scan,
                                    const uint32_t* __restrict__ pos_in, uint64_t count, uint32_t* __restrict__ pos_out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride)
    if (flag[j]) pos_out[scan[j]] = pos_in ? pos_in[j] : (uint32_t)j;
}

}  // namespace

int build_sa_bwt_device(const uint8_t* d_text, uint64_t n, uint32_t stride, cudaStream_t stream,
                        uint8_t** d_bwt_out, uint32_t** d_ssa_out, uint64_t* nsamp_out,
                        uint32_t** d_sa_out, uint32_t* rounds_out, uint32_t* passes_out, uint64_t* pair_passes_out) {
  uint32_t rounds = 0, passes = 0;
  uint64_t pair_passes = 0;
  if (rounds_out) *rounds_out = 0;
  if (passes_out) *passes_out = 0;
  if (pair_passes_out) *pair_passes_out = 0;
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

  PhaseTimer pt("suffix array");
  // 8-byte slack on the key buffers: the sparse rounds carve them into halves
  SA_CUDA(cudaMalloc(&key_a, n * 8 + 64));
  SA_CUDA(cudaMalloc(&key_b, n * 8 + 64));
  SA_CUDA(cudaMalloc(&sa_a, n * 4 + 64));
  SA_CUDA(cudaMalloc(&sa_b, n * 4 + 64));
  SA_CUDA(cudaMalloc(&rank, n * 4));
  SA_CUDA(cudaMalloc(&heads, n * 4 + 64));
  SA_CUDA(cudaMalloc(&d_present, 256 * 4));
  SA_CUDA(cudaMalloc(&d_ngroups, 16));

  pt.mark("allocate 32 n bytes");
  int dev = 0, num_sms = 148;
  SA_CUDA(cudaGetDevice(&dev));
  SA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));

  // alphabet
  SA_CUDA(cudaMemsetAsync(d_present, 0, 256 * 4, stream));
  byte_hist_kernel<<<num_sms * 8, 256, 0, stream>>>(d_text, n, d_present);
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

  const int grid = num_sms * 8, block = 256;
  init_keys_kernel<<<grid, block, 0, stream>>>(d_text, n, bps, k0, st, key_a, sa_a);

  cub::DoubleBuffer<uint64_t> keys(key_a, key_b);
  cub::DoubleBuffer<uint32_t> vals(sa_a, sa_b);
  size_t tmp_sort = 0, tmp_scan = 0, tmp_scan8 = 0;
  SA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, vals, (int64_t)n, 0, 64, stream));
  SA_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, heads, heads, MaxOp(), (int64_t)n, stream));
  SA_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_scan8, (const uint8_t*)nullptr, heads, cub::Sum(), (uint32_t)0, (int64_t)n, stream));
  const size_t tmp_bytes = std::max(std::max(tmp_sort, tmp_scan), tmp_scan8) + 256;
  SA_CUDA(cudaMalloc(&d_tmp, tmp_bytes));

  pt.mark("alphabet + packed keys");
  int nbits_n = 1;
  while ((1ull << nbits_n) < n + 1) ++nbits_n;
  uint64_t h = (uint64_t)k0;
  int end_bit = bps * k0;
  uint64_t u = 0;          // rows whose group has more than one suffix
  bool sparse = false;     // few enough of them left to sort only those
  // ---- dense rounds: every (key, suffix) pair is sorted --------------------------------------------------
  for (int round = 0;; ++round) {
    size_t tb = tmp_bytes;
    SA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys, vals, (int64_t)n, 0, end_bit, stream));
    ++rounds;
    passes += (uint32_t)((end_bit + 7) / 8);  // 8-bit digits: each pass reads and writes every (key, suffix) pair
    pair_passes += n * (uint64_t)((end_bit + 7) / 8);
    SA_CUDA(cudaMemsetAsync(d_ngroups, 0, 8, stream));
    mark_heads_kernel<<<grid, block, 0, stream>>>(keys.Current(), n, heads, d_ngroups);
    unsigned long long ngroups = 0;
    SA_CUDA(cudaMemcpyAsync(&ngroups, d_ngroups, 8, cudaMemcpyDeviceToHost, stream));
    SA_CUDA(cudaStreamSynchronize(stream));
    if (ngroups == n) break;
    tb = tmp_bytes;
    SA_CUDA(cub::DeviceScan::InclusiveScan(d_tmp, tb, heads, heads, MaxOp(), (int64_t)n, stream));
    scatter_rank_kernel<<<grid, block, 0, stream>>>(vals.Current(), heads, n, rank);
    // rows still unresolved: flags in the free key buffer, their exclusive scan in `heads` (free again)
    uint8_t* flag = reinterpret_cast<uint8_t*>(keys.Alternate());
    unresolved_flags_kernel<<<grid, block, 0, stream>>>(keys.Current(), n, flag);
    tb = tmp_bytes;
    SA_CUDA(cub::DeviceScan::ExclusiveScan(d_tmp, tb, flag, heads, cub::Sum(), (uint32_t)0, (int64_t)n, stream));
    uint32_t last_scan = 0;
    uint8_t last_flag = 0;
    SA_CUDA(cudaMemcpyAsync(&last_scan, heads + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
    SA_CUDA(cudaMemcpyAsync(&last_flag, flag + (n - 1), 1, cudaMemcpyDeviceToHost, stream));
    SA_CUDA(cudaStreamSynchronize(stream));
    u = (uint64_t)last_scan + last_flag;
    if (u * 2 <= n && !std::getenv("CSFM_SA_DENSE_ONLY")) {
      // compact the unresolved rows (in row order) into the free suffix buffer and go on with those only
      compact_rows_kernel<<<grid, block, 0, stream>>>(flag, heads, nullptr, n, vals.Alternate());
      sparse = true;
      break;
    }
    doubling_keys_kernel<<<grid, block, 0, stream>>>(vals.Current(), rank, n, h, keys.Current());
    h *= 2;
    end_bit = 32 + nbits_n;
    if (round > 40) {
      cleanup();
      return fail(CSFM_ERR_CUDA, "suffix sorting did not converge");
    }
  }
  pt.mark("dense rounds (sort all n pairs, group heads, ranks, unresolved rows)");
  uint32_t* const sa_final = vals.Current();
  if (sparse) {
    // ---- sparse rounds (u <= n / 2): all buffers are carved out of the dense ones -------------------------
    uint32_t* pos_a = vals.Alternate();               // u rows, then u more for the next round's
    uint32_t* pos_b = pos_a + u;
    uint64_t* const ku = key_a;                        // 2 x u keys
    uint32_t* const vu = heads;                        // 2 x u suffixes
    uint32_t* const heads_u = reinterpret_cast<uint32_t*>(key_b);        // u run heads / scan output
    uint8_t* const flag_u = reinterpret_cast<uint8_t*>(key_b) + 4 * n;   // u flags (u <= n / 2)
    end_bit = 32 + nbits_n;
    for (int round = 0; u > 0; ++round) {
      const int g = (int)std::min<uint64_t>((u + block - 1) / block, (uint64_t)grid);
      cub::DoubleBuffer<uint64_t> k2(ku, ku + u);
      cub::DoubleBuffer<uint32_t> v2(vu, vu + u);
      doubling_keys_sparse_kernel<<<g, block, 0, stream>>>(sa_final, pos_a, rank, n, u, h, k2.Current(), v2.Current());
      size_t tb = tmp_bytes;
      SA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, k2, v2, (int64_t)u, 0, end_bit, stream));
      ++rounds;
      pair_passes += u * (uint64_t)((end_bit + 7) / 8);
      sparse_heads_kernel<<<g, block, 0, stream>>>(k2.Current(), u, heads_u, flag_u);
      tb = tmp_bytes;
      SA_CUDA(cub::DeviceScan::InclusiveScan(d_tmp, tb, heads_u, heads_u, MaxOp(), (int64_t)u, stream));
      sparse_scatter_kernel<<<g, block, 0, stream>>>(v2.Current(), pos_a, heads_u, u, sa_final, rank);
      tb = tmp_bytes;
      SA_CUDA(cub::DeviceScan::ExclusiveScan(d_tmp, tb, flag_u, heads_u, cub::Sum(), (uint32_t)0, (int64_t)u, stream));
      uint32_t last_scan = 0;
      uint8_t last_flag = 0;
      SA_CUDA(cudaMemcpyAsync(&last_scan, heads_u + (u - 1), 4, cudaMemcpyDeviceToHost, stream));
      SA_CUDA(cudaMemcpyAsync(&last_flag, flag_u + (u - 1), 1, cudaMemcpyDeviceToHost, stream));
      SA_CUDA(cudaStreamSynchronize(stream));
      const uint64_t u_next = (uint64_t)last_scan + last_flag;
      if (u_next) compact_rows_kernel<<<g, block, 0, stream>>>(flag_u, heads_u, pos_a, u, pos_b);
      std::swap(pos_a, pos_b);
      u = u_next;
      h *= 2;
      if (round > 40) {
        cleanup();
        return fail(CSFM_ERR_CUDA, "suffix sorting did not converge");
      }
    }
  }
  pt.mark("sparse rounds");
  if (rounds_out) *rounds_out = rounds;
  if (passes_out) *passes_out = passes;
  if (pair_passes_out) *pair_passes_out = pair_passes;
  const uint32_t* d_sa = sa_final;
  SA_CUDA(cudaMalloc(&d_bwt, n));
  SA_CUDA(cudaMalloc(&d_ssa, nsamp * 4));
  bwt_ssa_kernel<<<grid, block, 0, stream>>>(d_text, d_sa, n, stride, d_bwt, d_ssa);
  SA_CUDA(cudaGetLastError());
  SA_CUDA(cudaStreamSynchronize(stream));
  if (d_sa_out) {  // hand the SA buffer over instead of freeing it
    if (d_sa == sa_a) { *d_sa_out = sa_a; sa_a = nullptr; } else { *d_sa_out = sa_b; sa_b = nullptr; }
  }
  pt.mark("BWT + samples");
  cleanup();
  pt.mark("free");
#undef SA_CUDA
  *d_bwt_out = d_bwt;
  *d_ssa_out = d_ssa;
  return CSFM_OK;
}

}  // namespace csfm
