// csfm_build.cu — BWT -> device-resident index (the "wm_build" kernels), both layouts.
//
// Replaces cs::WaveletTree::build + cs::BitVector::build
// (/root/reference/src/core/wavelet.cpp:14-53, src/core/bitvector.cpp:14-92) and the C-array
// loop of FMIndex::build_from_text (src/api/fm_index.cpp:36-47). Layouts: csfm_common.cuh.
//
// Layout 2 (nibble levels, 128-byte lines), per level:
//   nib_pack      one warp per 128-symbol line: packs the 16 payload words, per-line histogram
//                 of the 16 nibble values (match_any + shared counters)
//   cub::ExclusiveSum x16 over the per-line histograms -> absolute counters
//   nib_counters  writes the 16 counters of every line
//   nib_split     (level 0 of a 2-level index) stable 16-way partition by the high nibble:
//                 dst = start1[hi] + rank_0(hi, i) — the same formula the queries evaluate
// Layout 1 (binary, 64-byte lines), per level (bit L-1-l of the code):
//   pack_level / cub::ExclusiveSum / write_headers / split_level (stable 0/1 partition,
//   dst = bit ? zeros + rank1(i) : i - rank1(i), wavelet.cpp:47-51)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "csfm_dna.cuh"
#include "csfm_host.hpp"

namespace csfm {

namespace {

__global__ void histogram_kernel(const uint8_t* __restrict__ data, uint64_t n,
                                 unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
  for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n;
       base += stride) {
    if (base + 16 <= n) {
      const uint4 v = *reinterpret_cast<const uint4*>(data + base);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&sh[w[k] & 0xFF], 1u);
        atomicAdd(&sh[(w[k] >> 8) & 0xFF], 1u);
        atomicAdd(&sh[(w[k] >> 16) & 0xFF], 1u);
        atomicAdd(&sh[w[k] >> 24], 1u);
      }
    } else {
      for (uint64_t i = base; i < n; ++i) atomicAdd(&sh[data[i]], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

struct CodeTable {
  uint8_t code[256];
};

__global__ void map_codes_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                 uint64_t n, const __grid_constant__ CodeTable t) {
  __shared__ uint8_t sc[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sc[i] = t.code[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = sc[in[i]];
}

// ---- layout 1 -----------------------------------------------------------------------------

// One warp per line. Writes words 1..15 (payload) and 0 into word 0; line popcount to pop[].
__global__ void pack_level_kernel(const uint8_t* __restrict__ cur, uint64_t n, int bit,
                                  uint8_t* __restrict__ level, uint64_t nblk,
                                  uint32_t* __restrict__ pop) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t b = warp; b < nblk; b += nwarps) {
    const uint64_t i0 = b * kPayloadBits;
    uint32_t mine = 0, total = 0;
#pragma unroll
    for (int t = 0; t < 15; ++t) {
      const uint64_t i = i0 + 32 * t + lane;
      const uint32_t v = (i < n) ? ((cur[i] >> bit) & 1u) : 0u;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, v);
      if (lane == t + 1) mine = bal;  // lane k owns line word k
      total += __popc(bal);
    }
    if (lane < 16) reinterpret_cast<uint32_t*>(level + b * kLineBytes)[lane] = mine;
    if (lane == 0) pop[b] = total;
  }
}

__global__ void write_headers_kernel(uint8_t* __restrict__ level, uint64_t nblk,
                                     const uint32_t* __restrict__ rank_before) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += stride)
    *reinterpret_cast<uint32_t*>(level + b * kLineBytes) = rank_before[b];
}

// Stable 0/1 split of the symbol sequence for the next level.
__global__ void split_level_kernel(const uint8_t* __restrict__ cur, uint8_t* __restrict__ nxt,
                                   uint64_t n, int bit, uint64_t nblk,
                                   const uint32_t* __restrict__ rank_before, uint32_t zeros) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t b = warp; b < nblk; b += nwarps) {
    const uint64_t i0 = b * kPayloadBits;
    uint32_t r = rank_before[b];
#pragma unroll
    for (int t = 0; t < 15; ++t) {
      const uint64_t i = i0 + 32 * t + lane;
      const bool valid = i < n;
      const uint8_t sym = valid ? cur[i] : 0;
      const uint32_t v = valid ? ((sym >> bit) & 1u) : 0u;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, v);
      const uint32_t r1 = r + __popc(bal & lt);  // rank1(i)
      if (valid) {
        const uint64_t dst = v ? (uint64_t)zeros + r1 : i - r1;
        nxt[dst] = sym;
      }
      r += __popc(bal);
    }
  }
}

// ---- layout 2 -----------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 8;

// u32 index inside a 128-byte line of payload word w (0..15) / counter v (0..15)
__device__ __forceinline__ uint32_t line_word_of_payload(uint32_t w) { return 8u * (w >> 2) + 4u + (w & 3u); }
__device__ __forceinline__ uint32_t line_word_of_counter(uint32_t v) { return 8u * (v >> 2) + (v & 3u); }

// One warp per line: payload words + per-line histogram linecnt[v * nblk + b].
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
nib_pack_kernel(const uint8_t* __restrict__ cur, uint64_t n, int shift, uint8_t* __restrict__ level, uint64_t nblk,
                uint32_t* __restrict__ linecnt) {
  __shared__ uint32_t hist[kWarpsPerBlock][16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t b = warp; b < nblk; b += nwarps) {
    if (lane < 16) hist[wib][lane] = 0;
    __syncwarp();
    uint32_t* line = reinterpret_cast<uint32_t*>(level + b * kLine2Bytes);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint64_t i = b * kSymsPerLine + 32 * r + lane;
      const bool valid = i < n;
      const uint32_t v = valid ? ((cur[i] >> shift) & 15u) : 0u;
      // histogram: one shared add per distinct value per round
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? v : 0xFFu);
      if (valid && lane == __ffs(peers) - 1) hist[wib][v] += __popc(peers);
      // round r = chunk r, lane = symbol of the chunk; payload word b = bit b of the 32 symbols
      const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, v & 1u), b1 = __ballot_sync(0xFFFFFFFFu, v & 2u);
      const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, v & 4u), b3 = __ballot_sync(0xFFFFFFFFu, v & 8u);
      if (lane < 4) line[line_word_of_payload(4 * r + lane)] = lane == 0 ? b0 : lane == 1 ? b1 : lane == 2 ? b2 : b3;
      __syncwarp();
    }
    if (lane < 16) linecnt[(uint64_t)lane * nblk + b] = hist[wib][lane];
    __syncwarp();
  }
}

__global__ void nib_counters_kernel(uint8_t* __restrict__ level, uint64_t nblk, const uint32_t* __restrict__ prefix) {
  const uint64_t total = nblk * 16;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint64_t b = t >> 4;
    const uint32_t v = (uint32_t)(t & 15);
    reinterpret_cast<uint32_t*>(level + b * kLine2Bytes)[line_word_of_counter(v)] = prefix[(uint64_t)v * nblk + b];
  }
}

struct Start16 {
  uint32_t s[16];
};

// Stable 16-way partition by the high nibble: element i of level 0 goes to
// start1[hi] + rank_0(hi, i) in level 1.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
nib_split_kernel(const uint8_t* __restrict__ cur, uint8_t* __restrict__ nxt, uint64_t n, uint64_t nblk,
                 const uint32_t* __restrict__ prefix, const __grid_constant__ Start16 st) {
  __shared__ uint32_t run[kWarpsPerBlock][16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t b = warp; b < nblk; b += nwarps) {
    if (lane < 16) run[wib][lane] = st.s[lane] + prefix[(uint64_t)lane * nblk + b];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint64_t i = b * kSymsPerLine + 32 * r + lane;
      const bool valid = i < n;
      const uint8_t sym = valid ? cur[i] : 0;
      const uint32_t hi = sym >> 4;
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? hi : 0xFFu);
      if (valid) nxt[(uint64_t)run[wib][hi] + __popc(peers & lt)] = sym;
      __syncwarp();
      if (valid && lane == __ffs(peers) - 1) run[wib][hi] += __popc(peers);
      __syncwarp();
    }
  }
}

// ---- layout 3 (csfm_dna.cuh) --------------------------------------------------------------
// u32 index inside a 64-byte line of pair t (0..5) / counter v (0..3): lane h = t / 3 holds pairs 3h..3h+2
__device__ __forceinline__ uint32_t dna_word_of_pair(uint32_t t) { return 8u * (t / 3u) + 2u + 2u * (t % 3u); }
__device__ __forceinline__ uint32_t dna_word_of_counter(uint32_t v) { return 8u * (v >> 1) + (v & 1u); }

// the marked form (csfm_dna.cuh): pairs 0..3, two per half, then the half's two mark words; counters c0 c1 | c2 marks
__device__ __forceinline__ uint32_t dna_word_of_pair_m(uint32_t t) { return 8u * (t >> 1) + 2u + 2u * (t & 1u); }
__device__ __forceinline__ uint32_t dna_word_of_marks_m(uint32_t t) { return 8u * (t >> 1) + 6u + (t & 1u); }
__device__ __forceinline__ uint32_t dna_word_of_counter_m(uint32_t v) { return v < 2u ? v : 6u + v; }  // 0, 1, 8, 9

// One warp per line: the (lo, hi) pairs + the per-line histogram linecnt[v * nblk + b]. `cur` holds compact
// codes; the symbol that occurs once (code 4) is stored and counted as a 0. Marked form (kM): 128 rows per line, a
// mark bit for every row whose suffix starts at a multiple of the sample stride, and their number in slot v = 3
// (the counter of v = 3 is implied there).
template <bool kM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
dna_pack_kernel(const uint8_t* __restrict__ cur, uint64_t n, uint8_t* __restrict__ level, uint64_t nblk,
                uint32_t* __restrict__ linecnt, const uint32_t* __restrict__ sa, uint32_t stride) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t b = warp; b < nblk; b += nwarps) {
    uint32_t* line = reinterpret_cast<uint32_t*>(level + b * kLine3Bytes);
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (uint32_t t = 0; t < (kM ? 4u : 6u); ++t) {
      const uint64_t i = b * Line3<kM>::kSyms + 32 * t + lane;
      const bool valid = i < n;
      uint32_t v = valid ? cur[i] : 0u;
      if (v >= 4u) v = 0u;
      const uint32_t lo = __ballot_sync(0xFFFFFFFFu, v & 1u), hi = __ballot_sync(0xFFFFFFFFu, v & 2u);
      const uint32_t in = __ballot_sync(0xFFFFFFFFu, valid);
      c0 += __popc(~lo & ~hi & in);
      c1 += __popc(lo & ~hi & in);
      c2 += __popc(~lo & hi & in);
      if constexpr (kM) {
        const uint32_t mk = __ballot_sync(0xFFFFFFFFu, valid && sa[i] % stride == 0u);
        c3 += __popc(mk);
        if (lane == 0) {
          line[dna_word_of_pair_m(t)] = lo;
          line[dna_word_of_pair_m(t) + 1] = hi;
          line[dna_word_of_marks_m(t)] = mk;
        }
      } else {
        c3 += __popc(lo & hi & in);
        if (lane == 0) {
          line[dna_word_of_pair(t)] = lo;
          line[dna_word_of_pair(t) + 1] = hi;
        }
      }
    }
    if (lane < 4) linecnt[(uint64_t)lane * nblk + b] = lane == 0 ? c0 : lane == 1 ? c1 : lane == 2 ? c2 : c3;
  }
}

template <bool kM>
__global__ void dna_counters_kernel(uint8_t* __restrict__ level, uint64_t nblk, const uint32_t* __restrict__ prefix) {
  const uint64_t total = nblk * 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint64_t b = t >> 2;
    const uint32_t v = (uint32_t)(t & 3);
    reinterpret_cast<uint32_t*>(level + b * kLine3Bytes)[kM ? dna_word_of_counter_m(v) : dna_word_of_counter(v)] = prefix[(uint64_t)v * nblk + b];
  }
}

// the position samples of the marked form: the SA values that are multiples of the stride, in row order
struct IsPositionSample {
  uint32_t stride;
  __device__ __forceinline__ bool operator()(const uint32_t& s) const { return s % stride == 0u; }
};

// row of the one occurrence of `byte` in the BWT
__global__ void find_byte_kernel(const uint8_t* __restrict__ data, uint64_t n, uint8_t byte, unsigned int* __restrict__ row) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (data[i] == byte) *row = (unsigned int)i;
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

uint32_t bit_reverse(uint32_t v, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
}

}  // namespace

// Host-side derivation of every table from the byte histogram.
static void fill_tables(BlobHeader& h, const unsigned long long hist[256], uint32_t flags) {
  uint32_t cum = 0;
  for (int c = 0; c < 256; ++c) {  // fm_index.cpp:41-46
    h.C[c] = cum;
    cum += (uint32_t)hist[c];
  }
  h.C[256] = cum;
  uint32_t sigma = 0;
  std::memset(h.code_of_byte, 0, 256);
  std::memset(h.byte_of_code, 0, 256);
  const bool compact = !(flags & CSFM_BUILD_NO_COMPACT);
  for (int c = 0; c < 256; ++c) {
    if (hist[c]) {
      if (compact) {
        h.code_of_byte[c] = (uint8_t)sigma;
        h.byte_of_code[sigma] = (uint8_t)c;
      }
      ++sigma;
    }
  }
  h.sigma = sigma;
  uint32_t B = 8;
  if (compact) {
    B = 1;
    while ((1u << B) < sigma) ++B;
  } else {
    for (int c = 0; c < 256; ++c) {
      h.code_of_byte[c] = (uint8_t)c;
      h.byte_of_code[c] = (uint8_t)c;
    }
  }
  h.code_bits = B;
  std::vector<uint32_t> code_freq(256, 0);
  for (int c = 0; c < 256; ++c)
    if (hist[c]) code_freq[h.code_of_byte[c]] += (uint32_t)hist[c];
  std::memset(h.base_by_byte, 0, sizeof h.base_by_byte);
  std::memset(h.base_by_code, 0, sizeof h.base_by_code);
  std::memset(h.start1, 0, sizeof h.start1);

  if (h.layout == kLayoutDna64) {
    // up to four two-bit codes in byte order; with five symbols the one that occurs once (the smallest such
    // byte) becomes kSpecialCode and has no slot in the lines
    int special = -1;
    if (sigma == 5)
      for (int c = 0; c < 256 && special < 0; ++c)
        if (hist[c] == 1) special = c;
    std::memset(h.code_of_byte, 0, 256);
    std::memset(h.byte_of_code, 0, 256);
    uint32_t next = 0;
    for (int c = 0; c < 256; ++c) {
      if (!hist[c]) continue;
      const uint32_t code = c == special ? kSpecialCode : next++;
      h.code_of_byte[c] = (uint8_t)code;
      h.byte_of_code[code] = (uint8_t)c;
      h.base_by_byte[c] = h.C[c];  // one level: rank gives occ(c, .) itself
      h.base_by_code[code] = h.C[c];
    }
    h.levels = 1;
    h.code_bits = 2;
    h.special_byte = special < 0 ? 0u : (uint32_t)special;
    h.special_row = kNoSpecialRow;  // filled in once the BWT has been scanned
    return;
  }
  if (h.layout == kLayoutNibble128) {
    h.levels = B <= 4 ? 1 : 2;
    if (h.levels == 2) {
      uint32_t acc = 0;
      for (uint32_t g = 0; g < 16; ++g) {
        h.start1[g] = acc;
        for (uint32_t lo = 0; lo < 16; ++lo) acc += code_freq[(g << 4) | lo];
      }
    }
    for (int c = 0; c < 256; ++c) {
      if (!hist[c]) continue;
      const uint32_t code = h.code_of_byte[c];
      uint32_t before = 0;  // rank_1(lo, start1[hi]) = #{codes with a smaller hi and the same lo}
      if (h.levels == 2)
        for (uint32_t g = 0; g < (code >> 4); ++g) before += code_freq[(g << 4) | (code & 15u)];
      h.base_by_byte[c] = h.C[c] - before;  // u32 wrap-around is intended
      h.base_by_code[code] = h.C[c] - before;
    }
    return;
  }

  // layout 1: path(code,0) = number of symbols whose L-bit-reversed code is smaller: after L stable
  // 0/1 splits the sequence is ordered by the bit-reversed code (last split most significant).
  const uint32_t L = B;
  h.levels = L;
  const uint32_t ncodes = 1u << L;
  std::vector<uint64_t> freq_by_rev(ncodes, 0);
  for (uint32_t code = 0; code < ncodes; ++code) freq_by_rev[bit_reverse(code, L)] = code_freq[code];
  std::vector<uint32_t> start_by_rev(ncodes, 0);
  uint32_t acc = 0;
  for (uint32_t r = 0; r < ncodes; ++r) {
    start_by_rev[r] = acc;
    acc += (uint32_t)freq_by_rev[r];
  }
  for (int c = 0; c < 256; ++c) {
    if (!hist[c]) continue;
    const uint32_t code = h.code_of_byte[c];
    const uint32_t node_start = start_by_rev[bit_reverse(code, L)];
    h.base_by_byte[c] = h.C[c] - node_start;  // u32 wrap-around is intended
    h.base_by_code[code] = h.C[c] - node_start;
  }
}

int index_finish_handle(csfm_index* idx) {
  cudaDeviceProp prop;
  CSFM_CUDA(cudaGetDeviceProperties(&prop, idx->device));
  idx->num_sms = prop.multiProcessorCount;
  if (const char* e = std::getenv("CSFM_PATTERN_STAGING")) idx->tma_staging = std::strcmp(e, "tma") == 0;
  const BlobHeader& h = idx->h;
  IndexView& v = idx->view;
  v.levels = idx->d_blob + h.off_levels;
  v.levels_last = v.levels + (uint64_t)(h.levels - 1) * h.level_stride;
  v.ssa = reinterpret_cast<const uint32_t*>(idx->d_blob + h.off_ssa);
  v.hdr = reinterpret_cast<const BlobHeader*>(idx->d_blob);
  v.level_stride = h.level_stride;
  v.n = (uint32_t)h.n;
  v.L = h.levels;
  v.stride = h.stride;
  v.nsamp = (uint32_t)h.nsamp;
  v.layout = h.layout;
  v.stride_shift = 32;
  if ((h.stride & (h.stride - 1)) == 0)
    for (v.stride_shift = 0; (1u << v.stride_shift) < h.stride; ++v.stride_shift) {}
  v.special_row = h.layout == kLayoutDna64 ? h.special_row : kNoSpecialRow;
  v.special_byte = h.special_byte;
  v.special_first = h.C[h.special_byte & 0xFFu];
  v.marked = (h.layout == kLayoutDna64 && h.marked && h.off_psamp) ? 1u : 0u;
  v.psamp = v.marked ? reinterpret_cast<const uint32_t*>(idx->d_blob + h.off_psamp) : nullptr;
  v.pad1 = 0;
  v.kmer = h.kmer_k ? reinterpret_cast<const uint2*>(idx->d_blob + h.off_kmer) : nullptr;
  v.kmer_k = h.kmer_k;
  v.kmer_radix = h.kmer_radix;
  v.kmer_entries = 0;
  if (h.kmer_k) {
    v.kmer_entries = 1;
    for (uint32_t i = 0; i < h.kmer_k; ++i) v.kmer_entries *= h.kmer_radix;
  }
  v.kmer_tiled = h.kmer_tiled;
  v.kmer_hi = (h.kmer_k && h.off_kmer_hi) ? reinterpret_cast<const uint2*>(idx->d_blob + h.off_kmer_hi) : nullptr;
  if (std::getenv("CSFM_NO_HALF_TABLE")) v.kmer_hi = nullptr;  // experiment knob: ignore a table that is present
  v.text = h.off_text ? idx->d_blob + h.off_text : nullptr;
  v.dense = h.off_text ? reinterpret_cast<const uint32_t*>(idx->d_blob + h.off_dense) : nullptr;
  v.dense_shift = h.dense_shift;
  // a verification costs two HBM fetches and three loop trips; a step costs `levels` fetches and
  // one trip: worth it from 3 characters left on a two-level index, from 8 on a one-level index
  v.verify_min = h.verify_min ? h.verify_min : (h.levels >= 2 ? 3u : 8u);
  if (const char* e = std::getenv("CSFM_VERIFY_MIN")) v.verify_min = (uint32_t)std::atoi(e);
  idx->no_sa_locate = std::getenv("CSFM_NO_SA_LOCATE") != nullptr;
  if (const char* e = std::getenv("CSFM_COUNT3_LANES")) idx->count3_lanes = std::atoi(e) == 1 ? 1 : 2;
  if (const char* e = std::getenv("CSFM_WALK3_LANES")) idx->walk3_lanes = std::atoi(e) == 1 ? 1 : 2;
  // measured slower than the one-pass sub-warp kernel (profiles/README.md): selectable for A/B runs only
  idx->no_two_pass = std::getenv("CSFM_TWO_PASS") == nullptr;
  v.refill_min = 8;
  v.refill_wait = 5;
  if (const char* e = std::getenv("CSFM_REFILL_MIN")) v.refill_min = (uint32_t)std::min(8, std::max(1, std::atoi(e)));
  if (const char* e = std::getenv("CSFM_REFILL_WAIT")) v.refill_wait = (uint32_t)std::max(1, std::atoi(e));
  if (std::getenv("CSFM_NO_TEXT_CHECK")) v.text = nullptr;  // experiment knob: ignore the sections
  if (std::getenv("CSFM_NO_KMER_TABLE")) {  // experiment knob: ignore a table that is present
    v.kmer = nullptr;
    v.kmer_hi = nullptr;
    v.kmer_k = 0;
  }
  for (int l = 0; l < (int)kMaxLevels; ++l) v.zeros[l] = h.zeros[l];
  if (!idx->stream) CSFM_CUDA(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
  if (!idx->ev0) CSFM_CUDA(cudaEventCreate(&idx->ev0));
  if (!idx->ev1) CSFM_CUDA(cudaEventCreate(&idx->ev1));
  if (!idx->d_counters) {
    CSFM_CUDA(cudaMalloc(&idx->d_counters, kCounterSlots * kCounterWords * sizeof(unsigned long long)));
    CSFM_CUDA(cudaMemset(idx->d_counters, 0, kCounterSlots * kCounterWords * sizeof(unsigned long long)));
  }
  if (!idx->h_pinned) {
    CSFM_CUDA(cudaHostAlloc(&idx->h_pinned, 4096, cudaHostAllocMapped));
    std::memset(idx->h_pinned, 0, 4096);
    CSFM_CUDA(cudaHostGetDevicePointer(&idx->d_pinned, idx->h_pinned, 0));
  }
  return CSFM_OK;
}

int index_from_device_bwt(const uint8_t* d_bwt, uint64_t n, const uint32_t* d_ssa, uint64_t nsamp,
                          uint32_t stride, int device, uint32_t flags, csfm_index** out,
                          const uint8_t* d_text, const uint32_t* d_sa) {
  if (n > kMaxN) return fail(CSFM_ERR_TOO_LARGE, "text length must be < 2^32 - 1");
  if (stride == 0) return fail(CSFM_ERR_INVALID, "ssa_stride must be > 0");
  cudaStream_t st = nullptr;  // construction runs on the legacy default stream of the device
  PhaseTimer pt("index");

  // 1) byte histogram -> C, compact codes, level count, per-symbol base
  unsigned long long* d_hist = nullptr;
  CSFM_CUDA(cudaMalloc(&d_hist, 256 * sizeof(unsigned long long)));
  CSFM_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), st));
  if (n) histogram_kernel<<<1024, 256, 0, st>>>(d_bwt, n, d_hist);
  unsigned long long hist[256];
  CSFM_CUDA(cudaMemcpy(hist, d_hist, sizeof hist, cudaMemcpyDeviceToHost));
  cudaFree(d_hist);

  auto* idx = new csfm_index();
  idx->device = device;
  BlobHeader& h = idx->h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "CSFMDEV1", 8);
  h.version = 3;  // 3: bit-sliced chunk payload
  h.n = n;
  h.stride = stride;
  h.nsamp = nsamp;
  h.layout = (flags & CSFM_BUILD_LAYOUT_BINARY64) ? kLayoutBinary64 : kLayoutNibble128;
  h.special_row = kNoSpecialRow;
  if (std::getenv("CSFM_FORCE_TEXT_CHECK")) flags |= CSFM_BUILD_FORCE_TEXT_CHECK;
  {
    // layout 3 (two-bit symbols, 64-byte lines) whenever the text allows it: at most four symbols, or five of
    // which one occurs exactly once; the text-verification sections and raw byte codes exist in layout 2 only
    uint32_t present = 0, singles = 0;
    for (int c = 0; c < 256; ++c) {
      present += hist[c] != 0;
      singles += hist[c] == 1;
    }
    const uint32_t keep2 = CSFM_BUILD_LAYOUT_BINARY64 | CSFM_BUILD_NO_COMPACT | CSFM_BUILD_LAYOUT_NIBBLE128 |
                           CSFM_BUILD_FORCE_TEXT_CHECK | CSFM_BUILD_LARGE_TABLE;
    if (!(flags & keep2) && n > 0 && (present <= 4 || (present == 5 && singles >= 1)) && !std::getenv("CSFM_NO_DNA_LAYOUT"))
      h.layout = kLayoutDna64;
  }
  fill_tables(h, hist, flags);
  const uint32_t L = h.levels;
  const bool nib = h.layout == kLayoutNibble128;
  const bool dna = h.layout == kLayoutDna64;
  // layout 3, marked form (csfm_dna.cuh): suffix-array samples by text position, found through a mark bit per row. Needs
  // the suffix array at hand and SA[LF(r)] = SA[r] - 1 for every row but the one of suffix 0, i.e. a text whose LAST byte
  // occurs exactly once (any byte value; with five symbols it is the one without a two-bit code); other texts keep the
  // reference's row-sampled walk (and its failure semantics).
  bool marked = false;
  if (dna && n >= 2 && d_text && d_sa && !(flags & CSFM_BUILD_ROW_SAMPLES) && !std::getenv("CSFM_NO_POSITION_SAMPLES")) {
    uint8_t last = 0;
    marked = cudaMemcpy(&last, d_text + (n - 1), 1, cudaMemcpyDeviceToHost) == cudaSuccess && hist[last] == 1;
  }
  h.marked = marked ? 1u : 0u;
  h.nblk = dna ? n / (marked ? kSymsPerLine3M : kSymsPerLine3) + 1 : nib ? n / kSymsPerLine + 1 : n / kPayloadBits + 1;
  h.off_levels = kHeaderBytes;
  h.level_stride = align_up(h.nblk * (nib ? kLine2Bytes : kLineBytes), 256);  // layouts 1 and 3: 64-byte lines
  h.off_ssa = h.off_levels + (uint64_t)L * h.level_stride;
  h.total_bytes = align_up(h.off_ssa + nsamp * 4, 256);
  if (marked) {
    h.off_psamp = h.total_bytes;
    h.total_bytes = align_up(h.off_psamp + nsamp * 4, 256);
  }
  // k-mer jump table: the first k steps of a query become one lookup. Budget: a quarter of the
  // level bytes, between 1 MiB and 1 GiB (k = 3 for a byte alphabet at n = 2^30, 9 for DNA+$ at 2^26,
  // 11 for DNA+$ at 4e9).
  // (decided first because it also selects the table format) Text sections: see below.
  const bool levels_in_hbm = (uint64_t)L * h.level_stride > (96ull << 20);
  if (flags & CSFM_BUILD_LARGE_TABLE) {
    // after a long key little is left to step through: verify from three characters on, whatever the level count
    flags |= CSFM_BUILD_FORCE_TEXT_CHECK;
    h.verify_min = 3;
  }
  bool text_sections = false;
  if (nib && n >= 2 && d_text && d_sa && !(flags & CSFM_BUILD_NO_TEXT_CHECK) &&
      ((levels_in_hbm && L == 2) || (flags & CSFM_BUILD_FORCE_TEXT_CHECK))) {
    uint8_t last = 0;
    const cudaError_t e2 = cudaMemcpy(&last, d_text + (n - 1), 1, cudaMemcpyDeviceToHost);
    text_sections = e2 == cudaSuccess && hist[last] == 1 && h.C[last] == 0;
  }
  if ((nib || dna) && n && !(flags & CSFM_BUILD_NO_KMER_TABLE)) {
    // with the text on board the table holds sp only (4 bytes per key) and is filled from the text's k-gram histogram
    const bool tiled = text_sections && !std::getenv("CSFM_BUILD_PAIR_TABLE");
    const uint64_t entry_bytes = tiled ? 4 : 8;
    // layout 3: keys are built from the two-bit codes only (a pattern that contains the symbol occurring once
    // starts from the C array instead), so the table has 4^k entries, not 5^k
    const uint64_t radix = (flags & CSFM_BUILD_NO_COMPACT) ? 256 : dna ? std::min<uint64_t>(h.sigma, 4) : h.sigma;
    // layout 2: a quarter of the level bytes. Layout 3: half a byte per text symbol (1.5 x its level bytes) — its
    // count kernel is bound by issue slots, every character the lookup covers is a tenth of a 20-mer's work:
    // measured on C2 k = 10 / 11 / 12 -> 7.4 / 8.1 / 8.7e9 q/s at 39 / 64 / 165 MB; k = 11 keeps the index at the size of the text
    uint64_t budget = std::min<uint64_t>(1024ull << 20, std::max<uint64_t>(1ull << 20, dna ? n / 2 : (uint64_t)L * h.level_stride / 4));
    if (flags & CSFM_BUILD_LARGE_TABLE) {
      // opt-in: spend device memory on the table so that the lookup itself leaves few rows and the
      // query goes straight to the text verification (k = 4 for a byte alphabet at n = 2^30: 34 GB;
      // k = 13 for DNA+$ at n = 2^26: 9.8 GB)
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
      budget = std::max<uint64_t>(budget, std::min<uint64_t>(40ull << 30, free_b / 3));
    }
    if (const char* e = std::getenv("CSFM_KMER_BUDGET_MB")) budget = (uint64_t)std::max(1, std::atoi(e)) << 20;
    uint32_t k = 0;
    uint64_t entries = 1;
    // a key that already leaves a quarter of a row on average gains nothing from another character
    // (the kernels build keys in 32 bits: at most 2^32 of them)
    while (radix >= 2 && entries * radix * entry_bytes <= budget && entries * radix <= (1ull << 32) && k < 16 && entries < 4 * n) {
      entries *= radix;
      ++k;
    }
    if (k >= 2) {
      h.kmer_k = k;
      h.kmer_radix = (uint32_t)radix;
      h.kmer_tiled = tiled ? 1u : 0u;
      h.off_kmer = h.total_bytes;
      h.total_bytes = align_up(h.off_kmer + (tiled ? (entries + 1) * 4 : entries * 8), 256);
      // half-step table: 16 entries per k-mer; kept when it is no larger than the levels themselves
      // (C3: k = 3, 2.1 GB beside 2.3 GB of levels) and the table is not already a large one
      const uint64_t hi_bytes = entries * 16 * 8;
      if (L == 2 && !(flags & CSFM_BUILD_LARGE_TABLE) && hi_bytes <= std::max<uint64_t>((uint64_t)L * h.level_stride, 32ull << 20) &&
          hi_bytes <= (4ull << 30) && !std::getenv("CSFM_BUILD_NO_HALF_TABLE")) {
        h.off_kmer_hi = h.total_bytes;
        h.total_bytes = align_up(h.off_kmer_hi + hi_bytes, 256);
      }
    }
  }
  // Text-verification shortcut: needs the text and its suffix array, and a text whose last byte
  // is unique and the smallest present (then row r <-> suffix SA[r] and LF(r) <-> SA[r]-1 cyclically,
  // so "keep stepping from a one-row interval" equals "compare with the text"; without such a
  // terminator the reference's BWT is not a rotation BWT and the shortcut would change results).
  // Worth it only when the levels live in HBM: stepping through L2-resident lines is cheaper than
  // the two HBM fetches (suffix-array entry, text) of a verification.
  // On a one-level index (sigma <= 16) a step is a single fetch and the plain kernel (32 registers,
  // full occupancy) wins unless patterns are long: measured on the 4e9-byte DNA text, 2.68e9 q/s
  // plain vs 2.37e9 with the verification variant that never fires. So: two levels, in HBM.
  if (text_sections) {
    h.dense_shift = 0;  // the full suffix array: no extra steps to reach a sampled row
    h.off_text = h.total_bytes;
    h.off_dense = align_up(h.off_text + n + 64, 256);
    h.total_bytes = align_up(h.off_dense + (((n - 1) >> h.dense_shift) + 1) * 4, 256);
  }

  idx->blob_bytes = h.total_bytes;
  idx->owns_blob = true;
  cudaError_t e = cudaMalloc(&idx->d_blob, idx->blob_bytes);
  if (e != cudaSuccess) {
    delete idx;
    return fail(CSFM_ERR_NOMEM, std::string("cudaMalloc(index blob): ") + cudaGetErrorString(e));
  }

  pt.mark("histogram, tables, blob allocation");
  // 2) levels
  uint8_t *d_cur = nullptr, *d_nxt = nullptr;
  uint32_t *d_pop = nullptr, *d_rank = nullptr;
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  auto cleanup = [&]() {
    cudaFree(d_cur); cudaFree(d_nxt); cudaFree(d_pop); cudaFree(d_rank); cudaFree(d_scan_tmp);
  };
  auto bail = [&](int code, const std::string& m) {
    cleanup();
    csfm_destroy(idx);
    return fail(code, m);
  };
#define BUILD_CUDA(expr)                                                         \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess)                                                       \
      return bail(_e == cudaErrorMemoryAllocation ? CSFM_ERR_NOMEM : CSFM_ERR_CUDA, \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

  const uint64_t nbuf = n ? n : 1;
  const uint64_t ncnt = h.nblk * (nib ? 16 : dna ? 4 : 1);  // per-line counters before / after the scan
  BUILD_CUDA(cudaMalloc(&d_cur, nbuf));
  if (L > 1) BUILD_CUDA(cudaMalloc(&d_nxt, nbuf));
  BUILD_CUDA(cudaMalloc(&d_pop, ncnt * 4));
  BUILD_CUDA(cudaMalloc(&d_rank, ncnt * 4));
  BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp_bytes, d_pop, d_rank, (int64_t)h.nblk, st));
  BUILD_CUDA(cudaMalloc(&d_scan_tmp, scan_tmp_bytes ? scan_tmp_bytes : 16));

  if (n) {
    CodeTable ct;
    std::memcpy(ct.code, h.code_of_byte, 256);
    map_codes_kernel<<<2048, 256, 0, st>>>(d_bwt, d_cur, n, ct);
  }
  const int wpb = kWarpsPerBlock;
  const uint64_t want_blocks = (h.nblk + wpb - 1) / wpb;
  const int grid = (int)std::min<uint64_t>(want_blocks, 148ull * 64);
  if (dna) {
    uint8_t* level = idx->d_blob + h.off_levels;
    if (h.marked) dna_pack_kernel<true><<<grid, wpb * 32, 0, st>>>(d_cur, n, level, h.nblk, d_pop, d_sa, h.stride);
    else dna_pack_kernel<false><<<grid, wpb * 32, 0, st>>>(d_cur, n, level, h.nblk, d_pop, nullptr, 1u);
    for (int v = 0; v < 4; ++v)
      BUILD_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_tmp_bytes, d_pop + (uint64_t)v * h.nblk,
                                               d_rank + (uint64_t)v * h.nblk, (int64_t)h.nblk, st));
    if (h.marked) dna_counters_kernel<true><<<2048, 256, 0, st>>>(level, h.nblk, d_rank);
    else dna_counters_kernel<false><<<2048, 256, 0, st>>>(level, h.nblk, d_rank);
    if (h.marked) {
      // position samples: SA values that are multiples of the stride, in row order (d_pop is free again: its first
      // 8 bytes take the number selected)
      void* d_sel_tmp = nullptr;
      size_t sel_bytes = 0;
      uint32_t* const psamp = reinterpret_cast<uint32_t*>(idx->d_blob + h.off_psamp);
      unsigned long long* const d_nsel = reinterpret_cast<unsigned long long*>(d_pop);
      const IsPositionSample pred{h.stride};
      BUILD_CUDA(cub::DeviceSelect::If(nullptr, sel_bytes, d_sa, psamp, d_nsel, (int64_t)n, pred, st));
      BUILD_CUDA(cudaMalloc(&d_sel_tmp, sel_bytes ? sel_bytes : 16));
      cudaError_t se = cub::DeviceSelect::If(d_sel_tmp, sel_bytes, d_sa, psamp, d_nsel, (int64_t)n, pred, st);
      unsigned long long nsel = 0;
      if (se == cudaSuccess) se = cudaMemcpyAsync(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost, st);
      if (se == cudaSuccess) se = cudaStreamSynchronize(st);
      cudaFree(d_sel_tmp);
      BUILD_CUDA(se);
      if (nsel != nsamp) return bail(CSFM_ERR_CUDA, "position samples: the suffix array is not a permutation of 0..n-1");
    }
    if (h.code_of_byte[h.special_byte & 0xFFu] == kSpecialCode) {
      // the row of the symbol that occurs once (d_pop is free again: borrow its first word)
      BUILD_CUDA(cudaMemsetAsync(d_pop, 0xFF, 4, st));
      find_byte_kernel<<<1024, 256, 0, st>>>(d_bwt, n, (uint8_t)h.special_byte, d_pop);
      BUILD_CUDA(cudaMemcpyAsync(&h.special_row, d_pop, 4, cudaMemcpyDeviceToHost, st));
      BUILD_CUDA(cudaStreamSynchronize(st));
    }
  } else if (nib) {
    Start16 s16;
    std::memcpy(s16.s, h.start1, sizeof s16.s);
    for (uint32_t l = 0; l < L; ++l) {
      const int shift = (L == 2 && l == 0) ? 4 : 0;
      uint8_t* level = idx->d_blob + h.off_levels + (uint64_t)l * h.level_stride;
      nib_pack_kernel<<<grid, wpb * 32, 0, st>>>(d_cur, n, shift, level, h.nblk, d_pop);
      for (int v = 0; v < 16; ++v)
        BUILD_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_tmp_bytes, d_pop + (uint64_t)v * h.nblk,
                                                 d_rank + (uint64_t)v * h.nblk, (int64_t)h.nblk, st));
      nib_counters_kernel<<<2048, 256, 0, st>>>(level, h.nblk, d_rank);
      if (l + 1 < L && n) {
        nib_split_kernel<<<grid, wpb * 32, 0, st>>>(d_cur, d_nxt, n, h.nblk, d_rank, s16);
        std::swap(d_cur, d_nxt);
      }
    }
  } else {
    for (uint32_t l = 0; l < L; ++l) {
      const int bit = (int)(L - 1 - l);
      uint8_t* level = idx->d_blob + h.off_levels + (uint64_t)l * h.level_stride;
      pack_level_kernel<<<grid, wpb * 32, 0, st>>>(d_cur, n, bit, level, h.nblk, d_pop);
      BUILD_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_tmp_bytes, d_pop, d_rank, (int64_t)h.nblk, st));
      write_headers_kernel<<<1024, 256, 0, st>>>(level, h.nblk, d_rank);
      uint32_t last_rank = 0, last_pop = 0;
      BUILD_CUDA(cudaMemcpyAsync(&last_rank, d_rank + (h.nblk - 1), 4, cudaMemcpyDeviceToHost, st));
      BUILD_CUDA(cudaMemcpyAsync(&last_pop, d_pop + (h.nblk - 1), 4, cudaMemcpyDeviceToHost, st));
      BUILD_CUDA(cudaStreamSynchronize(st));
      const uint32_t ones = last_rank + last_pop;
      h.zeros[l] = (uint32_t)n - ones;
      if (l + 1 < L && n) {
        split_level_kernel<<<grid, wpb * 32, 0, st>>>(d_cur, d_nxt, n, bit, h.nblk, d_rank, h.zeros[l]);
        std::swap(d_cur, d_nxt);
      }
    }
  }
  pt.mark("levels");
  // 3) SA samples (+ text and suffix array for the verification shortcut) + header
  if (nsamp)
    BUILD_CUDA(cudaMemcpyAsync(idx->d_blob + h.off_ssa, d_ssa, nsamp * 4, cudaMemcpyDeviceToDevice, st));
  if (h.off_text) {
    BUILD_CUDA(cudaMemcpyAsync(idx->d_blob + h.off_text, d_text, n, cudaMemcpyDeviceToDevice, st));
    BUILD_CUDA(cudaMemsetAsync(idx->d_blob + h.off_text + n, 0, h.off_dense - (h.off_text + n), st));
    BUILD_CUDA(cudaMemcpyAsync(idx->d_blob + h.off_dense, d_sa, n * 4, cudaMemcpyDeviceToDevice, st));  // dense_shift == 0
  }
  {
    std::vector<uint8_t> hdr(kHeaderBytes, 0);
    std::memcpy(hdr.data(), &h, sizeof h);
    BUILD_CUDA(cudaMemcpyAsync(idx->d_blob, hdr.data(), kHeaderBytes, cudaMemcpyHostToDevice, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
  }
  BUILD_CUDA(cudaGetLastError());
  cleanup();
#undef BUILD_CUDA
  pt.mark("samples, text sections, header");
  int rc = index_finish_handle(idx);
  if (rc == CSFM_OK && h.kmer_k) rc = dna ? build_kmer_table3(idx, st) : build_kmer_table(idx, st);  // backward search over the finished levels
  pt.mark("handle + k-mer / half-step tables");
  if (rc != CSFM_OK) {
    csfm_destroy(idx);
    return rc;
  }
  *out = idx;
  return CSFM_OK;
}

}  // namespace csfm
