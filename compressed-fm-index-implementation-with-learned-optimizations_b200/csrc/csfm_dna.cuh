// csfm_dna.cuh — LAYOUT 3: two-bit symbols in 64-byte lines, for texts over at most four frequent
// symbols (DNA) plus at most one symbol that occurs exactly once (the terminator).
//
// Why a third layout. On B200 random fetches out of the L2 run at ~230 G 64-byte
// units/s but only ~145 G 128-byte units/s, and they keep that rate only while the working set stays
// under ~64-80 MB (profiles/r2_l2_sweep_probe.json). A DNA text in layout 2 spends a 4-bit slot and a
// 16-counter header on 5 symbols: 1 byte per symbol, 256 MB for the 2^28-byte text of configs[3]. Here a
// line holds 4 counters + 192 two-bit symbols: n/3 bytes (89 MB at 2^28, 22 MB at 2^26, 1.33 GB at 4e9),
// so C2 and C4 are served from the L2, and a line is fetched by a TWO-lane sub-warp: 16 queries per warp
// and half the warp instructions per rank.
//
//   line b (16 u32 words) = symbols [192 b, 192 b + 192):
//     lane 0 half: c0 c1 | lo0 hi0 | lo1 hi1 | lo2 hi2        lane 1 half: c2 c3 | lo3 hi3 | lo4 hi4 | lo5 hi5
//     c_v  = number of symbols == v before the line; pair t = symbols 32 t .. 32 t + 31 bit-sliced
//     (bit s of lo_t / hi_t = bit 0 / bit 1 of symbol 32 t + s).
//   rank(v, p) = line[p / 192].c_v + #{k < p % 192 : sym[k] == v}: one 64-byte line, two 256-bit loads.
//
// The symbol that occurs once (code 4, BWT row `special_row`) has no two-bit code: it is stored as a 0
// and counted as a 0, so rank(0, p) is one too high exactly when p > special_row — one compare fixes it —
// and occ(special, p) = (p > special_row) needs no memory at all.
//
// Replaces cs::WaveletTree::rank / access over 8 binary levels (/root/reference/src/core/wavelet.cpp:59-128)
// and cs::BitVector::rank1 (src/core/bitvector.cpp:165-230) for such texts; results are identical.
#pragma once
#include "csfm_common.cuh"

namespace csfm {

constexpr uint32_t kLayoutDna64 = 3;
constexpr uint32_t kSymsPerLine3 = 192;
constexpr uint32_t kLine3Bytes = 64;
constexpr uint32_t kSpecialCode = 4;          // compact code of the symbol that occurs once
constexpr uint32_t kNoSpecialRow = 0xFFFFFFFFu;

#ifdef __CUDACC__

// p / 192 and p % 192 without a divide: floor(x / 3) == umulhi(x, 0xAAAAAAAB) >> 1 for every 32-bit x
__device__ __forceinline__ uint32_t dna_line_of(uint32_t p) { return __umulhi(p >> 6, 0xAAAAAAABu) >> 1; }

// which of the lane's 96 symbols equal v: three hit words (pairs 0..2 of the lane's half)
struct DnaHits {
  uint32_t h0, h1, h2;
};
__device__ __forceinline__ DnaHits dna_hits(const Chunk32& k, uint32_t v) {
  // half line in a Chunk32: c0 = counter a, c1 = counter b, c2/c3 = pair 0, p0/p1 = pair 1, p2/p3 = pair 2
  const uint32_t m0 = bit_fill(v, 0), m1 = bit_fill(v, 1);
  DnaHits r;
  r.h0 = ~((k.c2 ^ m0) | (k.c3 ^ m1));
  r.h1 = ~((k.p0 ^ m0) | (k.p1 ^ m1));
  r.h2 = ~((k.p2 ^ m0) | (k.p3 ^ m1));
  return r;
}
// the lane's share of the counter of v: lane h holds the counters of v = 2h and 2h + 1
__device__ __forceinline__ uint32_t dna_counter(const Chunk32& k, uint32_t v, int h) {
  const uint32_t c = (v & 1u) ? k.c1 : k.c0;
  return ((v >> 1) == (uint32_t)h) ? c : 0u;
}
// partial rank of lane h for the in-line offset `off` (0..191): its counter share + hits below off
__device__ __forceinline__ uint32_t dna_partial(uint32_t cnt, const DnaHits& x, uint32_t off, int h) {
  const int loff = (int)off - 96 * h;
  return cnt + (uint32_t)__popc(x.h0 & low_mask(loff)) + (uint32_t)__popc(x.h1 & low_mask(loff - 32)) +
         (uint32_t)__popc(x.h2 & low_mask(loff - 64));
}
// the symbol at in-line offset off, valid in the lane that holds it (h == off >= 96)
__device__ __forceinline__ uint32_t dna_symbol(const Chunk32& k, uint32_t off, int h) {
  const uint32_t loff = off - 96u * (uint32_t)h;  // 0..95 in the owning lane
  const uint32_t t = loff >> 5, s = loff & 31u;
#ifdef CSFM_DNA_SYMBOL_BRANCHY
  const uint32_t lo = t == 0 ? k.c2 : (t == 1 ? k.p0 : k.p2);
  const uint32_t hi = t == 0 ? k.c3 : (t == 1 ? k.p1 : k.p3);
#else
  // selp chains (the ternary form compiles to divergent branches); in the other lane t is garbage: only its low bits are looked at
  const uint32_t lo = pick4(k.c2, k.p0, k.p2, k.p2, t);
  const uint32_t hi = pick4(k.c3, k.p1, k.p3, k.p3, t);
#endif
  return ((lo >> s) & 1u) | (((hi >> s) & 1u) << 1);
}
__device__ __forceinline__ uint32_t group2_sum(uint32_t v) { return v + __shfl_xor_sync(0xFFFFFFFFu, v, 1); }

// ---- one LANE per line (L2-resident indexes: count3_kernel<., 1>): the lane holds both halves A and B of the line ----
// the six hit words of the whole line for value v, and the counter of v
__device__ __forceinline__ void dna_line_hits(const Chunk32& A, const Chunk32& B, uint32_t v, uint32_t (&x)[6], uint32_t& cnt) {
  const uint32_t m0 = bit_fill(v, 0), m1 = bit_fill(v, 1);
  x[0] = ~((A.c2 ^ m0) | (A.c3 ^ m1));
  x[1] = ~((A.p0 ^ m0) | (A.p1 ^ m1));
  x[2] = ~((A.p2 ^ m0) | (A.p3 ^ m1));
  x[3] = ~((B.c2 ^ m0) | (B.c3 ^ m1));
  x[4] = ~((B.p0 ^ m0) | (B.p1 ^ m1));
  x[5] = ~((B.p2 ^ m0) | (B.p3 ^ m1));
  cnt = pick4(A.c0, A.c1, B.c0, B.c1, v);
}
// counter + hits below the in-line offset off (0..191)
__device__ __forceinline__ uint32_t dna_line_rank(uint32_t cnt, const uint32_t (&x)[6], uint32_t off) {
  uint32_t r = cnt;
#pragma unroll
  for (int t = 0; t < 6; ++t) r += (uint32_t)__popc(x[t] & low_mask((int)off - 32 * t));
  return r;
}

#endif  // __CUDACC__

}  // namespace csfm
