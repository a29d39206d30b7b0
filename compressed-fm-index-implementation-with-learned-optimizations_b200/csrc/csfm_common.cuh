// csfm_common.cuh — device-resident index layout and the rank primitive shared by all kernels.
//
// Layout (replaces the reference's two-level super/sub-block directory,
// /root/reference/src/core/bitvector.hpp:94-99, and its 8 std::vector-backed levels,
// src/core/wavelet.hpp:55-58):
//
//   level l of the wavelet matrix = nblk lines of 64 bytes, line b = 16 x u32:
//       word 0      : rank1 of the level before bit 480*b   (absolute, n < 2^32)
//       words 1..15 : bits [480*b, 480*b+480) of the level, LSB-first inside each word
//   nblk = n/480 + 1, so position p == n (ep starts at n) always has a line; bits past n are 0.
//   rank1_l(p) = line[p/480].word0 + popcount(payload bits below p%480): ONE 64-byte line,
//   fetched as 4 x 128-bit (or 2 x 256-bit) loads by a sub-warp — one sector pair in DRAM/L2.
//
// rank(c,i) of the reference (wavelet.cpp:59-96) follows TWO positions (start,end) per level;
// start always descends from 0, so it is a per-symbol constant:
//   path(c,i)   : p <- i; for each level, p <- bit ? zeros[l] + rank1_l(p) : p - rank1_l(p)
//   rank(c,i)   = path(c,i) - path(c,0)          (checked against the reference in tests)
// and backward search needs  C[c] + rank(c,i) = base[c] + path(c,i),  base[c] = C[c]-path(c,0).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace csfm {

constexpr uint32_t kPayloadBits = 480;  // bits per 64-byte line
constexpr uint32_t kLineBytes = 64;
constexpr uint32_t kMaxLevels = 8;
constexpr uint64_t kMaxN = 0xFFFFFFFEull;  // n < 2^32 - 1 (reference: u32 SA/C, fm_index.hpp:43-44)
constexpr uint32_t kAbsent = 0xFFFFFFFFu;

// Fixed-size header at the start of the device blob (also the host/.csidx representation).
struct BlobHeader {
  char magic[8];  // "CSFMDEV1"
  uint32_t version;
  uint32_t levels;  // L
  uint64_t n;
  uint32_t sigma;
  uint32_t stride;
  uint64_t nsamp;
  uint64_t nblk;          // lines per level
  uint64_t off_levels;    // byte offset of level 0
  uint64_t level_stride;  // bytes between levels (nblk*64 rounded up to 256)
  uint64_t off_ssa;       // byte offset of the SA samples (u32)
  uint64_t total_bytes;
  uint32_t zeros[kMaxLevels];  // number of 0 bits per level
  uint32_t reserved0[16];
  // byte-indexed tables
  uint32_t C[257];             // fm_index.cpp:36-47
  uint32_t pad0[3];
  uint32_t base_by_byte[256];  // C[b] - path(code(b),0)   (u32 wrap-around arithmetic)
  uint32_t base_by_code[256];  // same, indexed by compact code (LF step)
  uint8_t code_of_byte[256];   // compact code, 0 for absent bytes (absence <=> C[b+1]==C[b])
  uint8_t byte_of_code[256];
};
static_assert(sizeof(BlobHeader) % 16 == 0, "header must stay 16-byte aligned");
constexpr uint64_t kHeaderBytes = 4096;
static_assert(sizeof(BlobHeader) <= kHeaderBytes, "header grew past its slot");

// What a kernel needs, passed by value.
struct IndexView {
  const uint8_t* levels;  // level 0, line 0
  const uint32_t* ssa;
  const BlobHeader* hdr;  // device pointer (tables are staged to shared memory per CTA)
  uint64_t level_stride;
  uint32_t n;
  uint32_t L;
  uint32_t stride;
  uint32_t nsamp;
  uint32_t zeros[kMaxLevels];
};

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// mask of the low x bits, x clamped to [0,32]
__device__ __forceinline__ uint32_t low_mask(int x) {
  // bmsk.clamp would do; shifts with clamping semantics are as cheap
  return x <= 0 ? 0u : (x >= 32 ? 0xFFFFFFFFu : ((1u << x) - 1u));
}

// Partial rank of one lane of a 4-lane group. `w` = the lane's 16 bytes of the line (words
// 4j..4j+3), `off` = p % 480. Lane 0's word 0 is the absolute counter and is added as a value.
// Summing the four partials gives rank1(p).
__device__ __forceinline__ uint32_t lane_partial_rank(uint4 w, uint32_t off, int j) {
  // line word k covers payload bits [32(k-1), 32k): bits below off => x = off - 32(k-1)
  const int x0 = (int)off - 32 * (4 * j - 1);
  uint32_t r = (j == 0) ? w.x : (uint32_t)__popc(w.x & low_mask(x0));
  r += __popc(w.y & low_mask(x0 - 32));
  r += __popc(w.z & low_mask(x0 - 64));
  r += __popc(w.w & low_mask(x0 - 96));
  return r;
}

__device__ __forceinline__ uint32_t group4_sum(uint32_t v) {
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
  return v;
}

#endif  // __CUDACC__

}  // namespace csfm
