// csfm_common.cuh — device-resident index layouts and the rank primitives shared by all kernels.
//
// Both layouts replace the reference's two-level super/sub-block directory
// (/root/reference/src/core/bitvector.hpp:94-99) and its 8 std::vector-backed levels
// (src/core/wavelet.hpp:55-58). Symbols are first recoded to compact codes 0..sigma-1 (order
// preserving), B = ceil(log2 sigma) bits.
//
// LAYOUT 2 (default) — "nibble levels", 128-byte lines.  ncu on B200 shows that every L2 miss
// fills a whole 128-byte line from HBM and that random fetches top out at ~37 G lines/s whatever
// their size (profiles/README.md), so the unit of the index is the 128-byte line and each line
// resolves FOUR bits of the symbol instead of one:
//   levels  = 1 if B <= 4 (DNA, protein-reduced alphabets), else 2 (bytes): level 0 holds the
//             high part hi = code >> 4, level 1 the low nibbles lo = code & 15 stably grouped by
//             hi (a 16-ary wavelet matrix: group g starts at start1[g] = #{codes with hi < g}).
//   line b of a level = 128 symbols [128b, 128b+128) + 16 absolute counters, as 4 chunks of 32 B:
//       chunk j : cnt[4j..4j+3] (u32: #symbols == v in the level before this line)
//                 pay[0..3]     (u32: symbols 32j..32j+31 of the line, bit-sliced: word b holds
//                                bit b of each symbol, so the symbols equal to v come out as
//                                one hit word in symbol order after four logic operations)
//   rank_l(v,p) = line[p>>7].cnt[v] + #{k < (p&127) : sym[k] == v}: ONE line, fetched by a
//   4-lane sub-warp as four 256-bit loads (one L1 wavefront, four sectors, one DRAM fetch).
//   nblk = n/128 + 1 so that p == n (ep starts at n) has a line; padding symbols are 0 and are
//   never counted (they lie at or beyond every queried offset).
//   Backward search:  C[c] + occ(c,i) = base[c] + rank_last(lo, start1[hi] + rank_0(hi, i))
//   with base[c] = C[c] - rank_1(lo, start1[hi]) = C[c] - #{codes with smaller hi and the same lo}.
//
// LAYOUT 1 — binary wavelet matrix, 64-byte lines (the north star's literal layout; kept
// selectable with CSFM_BUILD_LAYOUT_BINARY64 for A/B measurements):
//   level l = nblk lines of 64 bytes: word 0 = rank1 before bit 480*b, words 1..15 = 480 bits.
//   path(c,i): p <- i; per level p <- bit ? zeros[l] + rank1_l(p) : p - rank1_l(p);
//   C[c] + occ(c,i) = base[c] + path(c,i), base[c] = C[c] - path(c,0).
//
// rank(c,i) of the reference (wavelet.cpp:59-96) follows TWO positions (start,end) per level;
// start always descends from 0, so it folds into the per-symbol constant base[c].
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace csfm {

constexpr uint32_t kLayoutBinary64 = 1;
constexpr uint32_t kLayoutNibble128 = 2;
// kLayoutDna64 = 3: csfm_dna.cuh

// layout 1
constexpr uint32_t kPayloadBits = 480;  // bits per 64-byte line
constexpr uint32_t kLineBytes = 64;
// layout 2
constexpr uint32_t kSymsPerLine = 128;  // 4-bit symbols per 128-byte line
constexpr uint32_t kLine2Bytes = 128;

constexpr uint32_t kMaxLevels = 8;
constexpr uint64_t kMaxN = 0xFFFFFFFEull;  // n < 2^32 - 1 (reference: u32 SA/C, fm_index.hpp:43-44)

// Fixed-size header at the start of the device blob (also the host/.csidx representation).
struct BlobHeader {
  char magic[8];  // "CSFMDEV1"
  uint32_t version;
  uint32_t levels;  // stored levels: layout 1: ceil(log2 sigma) or 8; layout 2: 1 or 2
  uint64_t n;
  uint32_t sigma;
  uint32_t stride;
  uint64_t nsamp;
  uint64_t nblk;          // lines per level
  uint64_t off_levels;    // byte offset of level 0
  uint64_t level_stride;  // bytes between levels (multiple of 256)
  uint64_t off_ssa;       // byte offset of the SA samples (u32)
  uint64_t total_bytes;
  uint32_t zeros[kMaxLevels];  // layout 1: number of 0 bits per level
  uint32_t layout;             // kLayoutBinary64 | kLayoutNibble128
  uint32_t code_bits;          // B
  // k-mer jump table (layout 2): entry e = (sp,ep) of the k-symbol string whose LAST symbol has
  // code e % radix, the one before it (e / radix) % radix, ...; (0,0) when it does not occur.
  uint32_t kmer_k;             // 0 = no table
  uint32_t kmer_radix;         // number of compact codes (sigma, or 256 with NO_COMPACT)
  uint64_t off_kmer;           // byte offset of the table (uint2 entries, or u32 when kmer_tiled)
  // text-verification shortcut (layout 2, texts with a unique smallest last byte): the text and
  // SA[k << dense_shift] for every k, so that a query whose interval has shrunk to ONE row can
  // finish by comparing its remaining characters with the text instead of stepping through them.
  uint64_t off_text;           // 0 = absent; n bytes + 64 bytes of padding, 16-byte aligned
  uint64_t off_dense;          // u32 samples
  uint32_t dense_shift;        // samples at rows that are multiples of 1 << dense_shift
  uint32_t verify_min;         // 0 = by level count (3 / 8); else: verify once at least this many characters are left
  // half-step table (two-level layout 2): for every k-mer e and every high nibble g, the interval
  // of the k-mer mapped through level 0 for g, i.e. (start1[g] + rank_0(g, sp), start1[g] + rank_0(g, ep)),
  // at entry e * 16 + g. A query with more than k characters starts from it and needs only the
  // level-1 half of its first rank step. 0 = absent.
  uint64_t off_kmer_hi;
  // 1: the k-mer table holds sp only, entries + 1 u32 values, ep(e) = sp(e + 1). Used when the text
  // ends in a unique smallest byte: rows are then sorted rotations, the intervals of all k-mers tile
  // [0, n) in key order, and the table is the prefix sum of the text's (cyclic) k-gram histogram.
  uint32_t kmer_tiled;
  // layout 3: 1 = the MARKED line form (csfm_dna.cuh: 128 rows per line with a mark bit each, suffix array sampled by text
  // position at off_psamp, nsamp entries) beside the row-sampled array of the reference's format at off_ssa
  uint32_t marked;
  uint32_t start1[16];         // layout 2: first position of hi-group g in level 1
  // byte-indexed tables
  uint32_t C[257];             // fm_index.cpp:36-47
  // layout 3 (csfm_dna.cuh): BWT row of the symbol that occurs once (0xFFFFFFFF: the text has none) and its byte
  uint32_t special_row;
  uint32_t special_byte;
  uint32_t pad0[1];
  uint32_t base_by_byte[256];  // see above (u32 wrap-around arithmetic)
  uint32_t base_by_code[256];  // same, indexed by compact code (LF step)
  uint8_t code_of_byte[256];   // compact code, 0 for absent bytes (absence <=> C[b+1]==C[b])
  uint8_t byte_of_code[256];
  uint64_t off_psamp;          // marked form: psamp[k] = SA value of the k-th marked row (u32), 0 = absent
  uint64_t reserved1;
};
static_assert(sizeof(BlobHeader) % 16 == 0, "header must stay 16-byte aligned");
constexpr uint64_t kHeaderBytes = 4096;
static_assert(sizeof(BlobHeader) <= kHeaderBytes, "header grew past its slot");

// What a kernel needs, passed by value.
struct IndexView {
  const uint8_t* levels;       // level 0, line 0
  const uint8_t* levels_last;  // last level, line 0 (== levels for a one-level index)
  const uint32_t* ssa;
  const BlobHeader* hdr;  // device pointer (tables are staged to shared memory per CTA)
  const uint2* kmer;      // k-mer jump table or nullptr
  const uint2* kmer_hi;   // half-step table or nullptr
  const uint8_t* text;    // text for the verification shortcut or nullptr
  const uint32_t* dense;  // SA[k << dense_shift]
  uint32_t kmer_k;
  uint32_t kmer_radix;
  uint32_t kmer_tiled;  // table format, see BlobHeader
  uint32_t special_row;  // layout 3: BWT row of the symbol that occurs once, 0xFFFFFFFF = none
  uint64_t level_stride;
  uint32_t n;
  uint32_t L;
  uint32_t stride;
  uint32_t nsamp;
  uint32_t layout;
  uint32_t stride_shift;  // log2(stride) when the stride is a power of two, else 32
  uint32_t dense_shift;
  uint32_t verify_min;  // verify against the text only when at least this many characters are left
  uint32_t refill_min;   // shortcut kernel: a warp fetches new queries once this many of its sub-warps are idle ...
  uint32_t refill_wait;  // ... or this many trips after its last refill, whichever comes first
  uint32_t special_first;  // layout 3: C[special byte] = the row LF maps special_row to
  uint32_t special_byte;
  uint64_t kmer_entries;   // keys of the k-mer table (radix ^ k), 0 = no table
  const uint32_t* psamp;   // layout 3, marked form: SA values of the marked rows in row order, else nullptr
  uint32_t marked;         // layout 3: the lines are in the marked form
  uint32_t pad1;
  uint32_t zeros[kMaxLevels];
};

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__

// Debug build (-DCSFM_BOUNDS_CHECK, tools/bounds_check.py: the stand-in for compute-sanitizer, which is closed on the GPU
// pool): every data-dependent address the query kernels form — level lines, table entries, suffix-array entries, text
// windows, pattern bytes — is checked against the extent of its section before the load; a violation prints and
// traps, so the launch (and the test) fails instead of reading a neighbour's bytes. Compiles to nothing otherwise.
#ifdef CSFM_BOUNDS_CHECK
#define CSFM_CHK(cond, what)                                                                     \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      printf("CSFM_BOUNDS_CHECK: %s violated at %s:%d (block %d thread %d)\n", what, __FILE__, __LINE__, \
             (int)blockIdx.x, (int)threadIdx.x);                                                 \
      __trap();                                                                                  \
    }                                                                                            \
  } while (0)
#else
#define CSFM_CHK(cond, what) ((void)0)
#endif

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

struct Chunk32 {  // one 32-byte chunk of a layout-2 line: 4 counters + 4 payload words
  uint32_t c0, c1, c2, c3, p0, p1, p2, p3;
};

// 256-bit load (LDG.E.256 on sm_100a): one instruction per lane, one wavefront per 4-lane group.
__device__ __forceinline__ Chunk32 ldg_nc_v8(const void* p) {
  Chunk32 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.c0), "=r"(r.c1), "=r"(r.c2), "=r"(r.c3), "=r"(r.p0), "=r"(r.p1), "=r"(r.p2), "=r"(r.p3)
               : "l"(p));
  return r;
}

// L2 residency hints for the layout-3 kernels (CSFM_L2_HINTS: 0 none, 1 one-shot data evict-first, 2 + index lines and
// table entries evict-last): what a walk or a search comes back to (level lines, k-mer table) should outlive what it
// touches once (suffix-array samples, result stores) in a 126 MB L2 that the index only just fits or just exceeds.
#ifndef CSFM_L2_HINTS
#define CSFM_L2_HINTS 0
#endif
__device__ __forceinline__ Chunk32 ldg_line_keep(const void* p) {
#if CSFM_L2_HINTS >= 2
  Chunk32 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.c0), "=r"(r.c1), "=r"(r.c2), "=r"(r.c3), "=r"(r.p0), "=r"(r.p1), "=r"(r.p2), "=r"(r.p3)
               : "l"(p));
  return r;
#else
  return ldg_nc_v8(p);
#endif
}
__device__ __forceinline__ uint2 ldg_table_keep(const uint2* p) {
#if CSFM_L2_HINTS >= 2
  uint64_t pol;
  uint2 r;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
#else
  return *p;
#endif
}
__device__ __forceinline__ uint32_t ldg_once_u32(const uint32_t* p) {
#if CSFM_L2_HINTS >= 1
  uint64_t pol;
  uint32_t r;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
#else
  return *p;
#endif
}
__device__ __forceinline__ void stg_once_u64(uint64_t* p, uint64_t v) {
#if CSFM_L2_HINTS >= 1
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
#else
  *p = v;
#endif
}

// A Chunk32 whose registers are "defined" for the compiler without costing an instruction. Lanes
// that skip the load (inactive sub-warp, or ep in the same line as sp) compute on garbage that is
// discarded: results only travel inside the 4-lane group and are committed under `if (active)`.
__device__ __forceinline__ Chunk32 chunk_undefined() {
  Chunk32 r;
  asm("" : "=r"(r.c0), "=r"(r.c1), "=r"(r.c2), "=r"(r.c3), "=r"(r.p0), "=r"(r.p1), "=r"(r.p2), "=r"(r.p3));
  return r;
}

// mask of the low x bits, x clamped to [0,32]:  high word of (0x00000000FFFFFFFF << min(x,32))
__device__ __forceinline__ uint32_t low_mask(int x) {
  return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(x, 0));
}

// layout 1: partial rank of one lane of a 4-lane group. `w` = the lane's 16 bytes of the line
// (words 4j..4j+3), `off` = p % 480. Lane 0's word 0 is the absolute counter (added as a value).
__device__ __forceinline__ uint32_t lane_partial_rank(uint4 w, uint32_t off, int j) {
  // line word k covers payload bits [32(k-1), 32k): bits below off => x = off - 32(k-1)
  const int x0 = (int)off - 32 * (4 * j - 1);
  uint32_t r = (j == 0) ? w.x : (uint32_t)__popc(w.x & low_mask(x0));
  r += __popc(w.y & low_mask(x0 - 32));
  r += __popc(w.z & low_mask(x0 - 64));
  r += __popc(w.w & low_mask(x0 - 96));
  return r;
}

// layout 2: the 32 symbols of a chunk are stored BIT-SLICED: payload word b holds bit b of every
// symbol (bit s of word b = bit b of symbol s). "Which symbols equal v" is then four logic
// operations on whole words and comes out in symbol order:
// bit s of chunk_hits() is set <=> symbol s of the chunk equals v (0..15).
__device__ __forceinline__ uint32_t bit_fill(uint32_t v, int b) {  // all ones if bit b of v is set, else 0
  return 0u - ((v >> b) & 1u);
}
__device__ __forceinline__ uint32_t chunk_hits(const Chunk32& k, uint32_t v) {
  uint32_t miss = k.p0 ^ bit_fill(v, 0);
  miss |= k.p1 ^ bit_fill(v, 1);
  miss |= k.p2 ^ bit_fill(v, 2);
  miss |= k.p3 ^ bit_fill(v, 3);
  return ~miss;
}
// Branch-free pick of one of four registers by a 2-bit index (selp chain: the ternary form
// compiles to divergent branches).
__device__ __forceinline__ uint32_t pick4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t idx) {
  uint32_t r;
  asm("{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .u32 t0, t1, b0, b1;\n\t"
      "and.b32 b0, %5, 1;\n\t"
      "and.b32 b1, %5, 2;\n\t"
      "setp.ne.u32 p0, b0, 0;\n\t"
      "setp.ne.u32 p1, b1, 0;\n\t"
      "selp.u32 t0, %2, %1, p0;\n\t"
      "selp.u32 t1, %4, %3, p0;\n\t"
      "selp.u32 %0, t1, t0, p1;\n\t"
      "}"
      : "=r"(r)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(idx));
  return r;
}
// cnt[v] if this lane's chunk holds it (lane j holds cnt[4j..4j+3]), else 0
__device__ __forceinline__ uint32_t chunk_counter(const Chunk32& k, uint32_t v, int j) {
  const uint32_t cnt = pick4(k.c0, k.c1, k.c2, k.c3, v);
  return ((v >> 2) == (uint32_t)j) ? cnt : 0u;
}
// partial rank of lane j for offset off = p & 127, given the lane's counter share and hit word
__device__ __forceinline__ uint32_t chunk_partial(uint32_t cnt, uint32_t hits, uint32_t off, int j) {
  return cnt + (uint32_t)__popc(hits & low_mask((int)off - 32 * j));
}
// the symbol at offset off of a line, given this lane's chunk (valid in lane off>>5 only)
__device__ __forceinline__ uint32_t chunk_symbol(const Chunk32& k, uint32_t off) {
  const uint32_t s = off & 31u;
  return ((k.p0 >> s) & 1u) | (((k.p1 >> s) & 1u) << 1) | (((k.p2 >> s) & 1u) << 2) | (((k.p3 >> s) & 1u) << 3);
}

// SA rows are sampled at multiples of the stride (fm_index.cpp:57-65): no hardware divide for the
// usual power-of-two strides.
__device__ __forceinline__ bool row_is_sampled(const IndexView& iv, uint32_t row) {
  return iv.stride_shift < 32 ? (row & (iv.stride - 1u)) == 0u : row % iv.stride == 0u;
}
__device__ __forceinline__ uint32_t sample_index(const IndexView& iv, uint32_t row) {
  return iv.stride_shift < 32 ? row >> iv.stride_shift : row / iv.stride;
}

// a `bytes`-byte load at `p` must lie inside the level lines of the index
__device__ __forceinline__ void check_line(const IndexView& iv, const void* p, uint32_t bytes) {
  CSFM_CHK(reinterpret_cast<const uint8_t*>(p) >= iv.levels &&
               reinterpret_cast<const uint8_t*>(p) + bytes <= iv.levels + (uint64_t)iv.L * iv.level_stride,
           "level line inside the index");
  (void)iv; (void)p; (void)bytes;
}

__device__ __forceinline__ uint32_t group4_sum(uint32_t v) {
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
  return v;
}

#endif  // __CUDACC__

}  // namespace csfm
