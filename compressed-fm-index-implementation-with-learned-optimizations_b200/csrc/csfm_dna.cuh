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

// ---- the MARKED form of the line (BlobHeader::marked; texts whose last byte is the symbol that occurs once) ----------
// locate needs SA[row]; the reference samples the suffix array by ROW (ssa[k] = SA[k * stride], fm_index.cpp:60-64) and
// walks LF until it meets a sampled row: a geometric walk, stride - 1 steps on average and unbounded in the tail (516
// on C4). Sampling by TEXT POSITION (rows whose SA value is a multiple of the stride) bounds the walk by stride - 1
// and halves its mean — SA[LF(r)] = SA[r] - 1 when the last byte of the text is unique — with the same number of
// samples, but a row must then carry a mark bit, and the walk reads it from the line it fetches anyway:
//
//   line b (16 u32 words) = rows [128 b, 128 b + 128):
//     lane 0 half: c0 c1 | lo0 hi0 | lo1 hi1 | mk0 mk1        lane 1 half: c2 mr | lo2 hi2 | lo3 hi3 | mk2 mk3
//     mk_t = mark bits of rows 32 t .. 32 t + 31; mr = number of marked rows before the line (the index into the
//     position samples); the counter of v = 3 is not stored: c3 = 128 b - c0 - c1 - c2.
//
// Every row-to-position answer is SA[row] either way, so results are identical; the row-sampled array of the
// reference's format stays in the blob for export (ssa(), .csidx) and for the text export.
constexpr uint32_t kSymsPerLine3M = 128;
template <bool kM> struct Line3 {
  static constexpr uint32_t kSyms = kM ? kSymsPerLine3M : kSymsPerLine3;
  static constexpr uint32_t kHalf = kSyms / 2;
};

// p / 192 and p % 192 without a divide: floor(x / 3) == umulhi(x, 0xAAAAAAAB) >> 1 for every 32-bit x
__device__ __forceinline__ uint32_t dna_line_of(uint32_t p) { return __umulhi(p >> 6, 0xAAAAAAABu) >> 1; }
template <bool kM> __device__ __forceinline__ uint32_t dna_line_of_t(uint32_t p) {
  if constexpr (kM) return p >> 7;
  else return dna_line_of(p);
}

// which of the lane's symbols equal v: three hit words (pairs 0..2 of the lane's half; two in the marked form)
struct DnaHits {
  uint32_t h0, h1, h2;
};
template <bool kM = false>
__device__ __forceinline__ DnaHits dna_hits(const Chunk32& k, uint32_t v) {
  // half line in a Chunk32: c0 = counter a, c1 = counter b, c2/c3 = pair 0, p0/p1 = pair 1, p2/p3 = pair 2 (marks in the marked form)
  const uint32_t m0 = bit_fill(v, 0), m1 = bit_fill(v, 1);
  DnaHits r;
  r.h0 = ~((k.c2 ^ m0) | (k.c3 ^ m1));
  r.h1 = ~((k.p0 ^ m0) | (k.p1 ^ m1));
  r.h2 = kM ? 0u : ~((k.p2 ^ m0) | (k.p3 ^ m1));
  return r;
}
// the lane's share of the counter of v: lane h holds the counters of v = 2h and 2h + 1. Marked form: lane 1 holds
// c2 and the mark counter; the shares of v = 3 add up to line_start - c0 - c1 - c2 (u32 wrap-around intended).
template <bool kM = false>
__device__ __forceinline__ uint32_t dna_counter(const Chunk32& k, uint32_t v, int h, uint32_t line_start = 0) {
  if constexpr (kM) {
    const uint32_t a = h == 0 ? (v == 0u ? k.c0 : v == 1u ? k.c1 : 0u) : (v == 2u ? k.c0 : 0u);
    const uint32_t three = h == 0 ? 0u - (k.c0 + k.c1) : line_start - k.c0;
    return v == 3u ? three : a;
  } else {
    const uint32_t c = (v & 1u) ? k.c1 : k.c0;
    return ((v >> 1) == (uint32_t)h) ? c : 0u;
  }
}
// partial rank of lane h for the in-line offset `off`: its counter share + hits below off
template <bool kM = false>
__device__ __forceinline__ uint32_t dna_partial(uint32_t cnt, const DnaHits& x, uint32_t off, int h) {
  const int loff = (int)off - (int)Line3<kM>::kHalf * h;
  uint32_t r = cnt + (uint32_t)__popc(x.h0 & low_mask(loff)) + (uint32_t)__popc(x.h1 & low_mask(loff - 32));
  if constexpr (!kM) r += (uint32_t)__popc(x.h2 & low_mask(loff - 64));
  return r;
}
// the symbol at in-line offset off, valid in the lane that holds it (h == off >= half)
template <bool kM = false>
__device__ __forceinline__ uint32_t dna_symbol(const Chunk32& k, uint32_t off, int h) {
  const uint32_t loff = off - Line3<kM>::kHalf * (uint32_t)h;  // 0 .. half - 1 in the owning lane
  const uint32_t t = loff >> 5, s = loff & 31u;
#ifdef CSFM_DNA_SYMBOL_BRANCHY
  const uint32_t lo = t == 0 ? k.c2 : (t == 1 ? k.p0 : k.p2);
  const uint32_t hi = t == 0 ? k.c3 : (t == 1 ? k.p1 : k.p3);
#else
  // selp chains (the ternary form compiles to divergent branches); in the other lane t is garbage: only its low bits are looked at
  const uint32_t lo = kM ? ((t & 1u) ? k.p0 : k.c2) : pick4(k.c2, k.p0, k.p2, k.p2, t);
  const uint32_t hi = kM ? ((t & 1u) ? k.p1 : k.c3) : pick4(k.c3, k.p1, k.p3, k.p3, t);
#endif
  return ((lo >> s) & 1u) | (((hi >> s) & 1u) << 1);
}
// marked form: the mark bit of the row at in-line offset off (valid in the owning lane) ...
__device__ __forceinline__ uint32_t dna_mark(const Chunk32& k, uint32_t off, int h) {
  const uint32_t loff = off - kSymsPerLine3M / 2 * (uint32_t)h;
  return (((loff & 32u) ? k.p3 : k.p2) >> (loff & 31u)) & 1u;
}
// ... and the lane's share of the number of marked rows before it (lane 1 adds the line's mark counter)
__device__ __forceinline__ uint32_t dna_mark_partial(const Chunk32& k, uint32_t off, int h) {
  const int loff = (int)off - (int)(kSymsPerLine3M / 2) * h;
  return (h == 1 ? k.c1 : 0u) + (uint32_t)__popc(k.p2 & low_mask(loff)) + (uint32_t)__popc(k.p3 & low_mask(loff - 32));
}
__device__ __forceinline__ uint32_t group2_sum(uint32_t v) { return v + __shfl_xor_sync(0xFFFFFFFFu, v, 1); }

// ---- one LANE per line (L2-resident indexes: count3_kernel<., 1>): the lane holds both halves A and B of the line ----
// the hit words of the whole line for value v (six; four in the marked form), and the counter of v
template <bool kM = false>
__device__ __forceinline__ void dna_line_hits(const Chunk32& A, const Chunk32& B, uint32_t v, uint32_t (&x)[6], uint32_t& cnt,
                                              uint32_t line_start = 0) {
  const uint32_t m0 = bit_fill(v, 0), m1 = bit_fill(v, 1);
  if constexpr (kM) {
    x[0] = ~((A.c2 ^ m0) | (A.c3 ^ m1));
    x[1] = ~((A.p0 ^ m0) | (A.p1 ^ m1));
    x[2] = ~((B.c2 ^ m0) | (B.c3 ^ m1));
    x[3] = ~((B.p0 ^ m0) | (B.p1 ^ m1));
    x[4] = x[5] = 0u;
    cnt = pick4(A.c0, A.c1, B.c0, line_start - A.c0 - A.c1 - B.c0, v);
  } else {
    x[0] = ~((A.c2 ^ m0) | (A.c3 ^ m1));
    x[1] = ~((A.p0 ^ m0) | (A.p1 ^ m1));
    x[2] = ~((A.p2 ^ m0) | (A.p3 ^ m1));
    x[3] = ~((B.c2 ^ m0) | (B.c3 ^ m1));
    x[4] = ~((B.p0 ^ m0) | (B.p1 ^ m1));
    x[5] = ~((B.p2 ^ m0) | (B.p3 ^ m1));
    cnt = pick4(A.c0, A.c1, B.c0, B.c1, v);
  }
}
// counter + hits below the in-line offset off
template <bool kM = false>
__device__ __forceinline__ uint32_t dna_line_rank(uint32_t cnt, const uint32_t (&x)[6], uint32_t off) {
  uint32_t r = cnt;
#pragma unroll
  for (int t = 0; t < (kM ? 4 : 6); ++t) r += (uint32_t)__popc(x[t] & low_mask((int)off - 32 * t));
  return r;
}
// the symbol at in-line offset off of a whole line
template <bool kM = false>
__device__ __forceinline__ uint32_t dna_line_symbol(const Chunk32& A, const Chunk32& B, uint32_t off) {
  uint32_t lo, hi;
  if constexpr (kM) {
    const uint32_t t = off >> 5;  // 0..3
    lo = pick4(A.c2, A.p0, B.c2, B.p0, t);
    hi = pick4(A.c3, A.p1, B.c3, B.p1, t);
  } else {
    const bool in_b = off >= 96u;  // half A holds symbols 0..95, half B 96..191
    const uint32_t t = (off - (in_b ? 96u : 0u)) >> 5;
    lo = in_b ? pick4(B.c2, B.p0, B.p2, B.p2, t) : pick4(A.c2, A.p0, A.p2, A.p2, t);
    hi = in_b ? pick4(B.c3, B.p1, B.p3, B.p3, t) : pick4(A.c3, A.p1, A.p3, A.p3, t);
  }
  const uint32_t s = off & 31u;  // 96 and 64 are multiples of 32
  return ((lo >> s) & 1u) | (((hi >> s) & 1u) << 1);
}
// marked form, whole line: the mark bit of the row at off, and the number of marked rows before it (index of its sample)
__device__ __forceinline__ uint32_t dna_line_mark(const Chunk32& A, const Chunk32& B, uint32_t off) {
  return (pick4(A.p2, A.p3, B.p2, B.p3, off >> 5) >> (off & 31u)) & 1u;
}
__device__ __forceinline__ uint32_t dna_line_mark_rank(const Chunk32& A, const Chunk32& B, uint32_t off) {
  return B.c1 + (uint32_t)__popc(A.p2 & low_mask((int)off)) + (uint32_t)__popc(A.p3 & low_mask((int)off - 32)) +
         (uint32_t)__popc(B.p2 & low_mask((int)off - 64)) + (uint32_t)__popc(B.p3 & low_mask((int)off - 96));
}

#endif  // __CUDACC__

}  // namespace csfm
