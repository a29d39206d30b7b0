// csfm_query3.cu — query kernels for layout 3 (two-bit symbols, 64-byte lines; csfm_dna.cuh), sm_100a.
//
// A TWO-lane sub-warp owns one query (count) / one occurrence row (locate): one rank = one 64-byte line =
// two 256-bit loads, three hit words and three masked popc per lane, one xor-shuffle. Sixteen queries per
// warp run one loop in lock-step; a loop trip is one backward-search step / one LF step; a sub-warp that
// finishes refills from a warp-local chunk of the batch cursor (ballot-ranked, one global atomic per 32
// queries). Grids are persistent. The first k characters of a query come from the k-mer table.
//
// Replaces cs::FMIndex::count / locate (/root/reference/src/api/fm_index.cpp:79-157) and
// cs::WaveletTree::rank / access (src/core/wavelet.cpp:59-128) for texts over <= 4 frequent symbols.
#include <algorithm>

#include "csfm_dna.cuh"
#include "csfm_host.hpp"
#include "csfm_kernels.cuh"

namespace csfm {

namespace {

// rank(v, sp) and rank(v, ep) in the one level of the index. lv = level base + 32 * h. All 32 lanes call it.
// v == kSpecialCode (the symbol that occurs once, at BWT row px): occ(v, p) = (p > px), no memory access.
template <bool kM>
__device__ __forceinline__ void rank_pair3(const uint8_t* __restrict__ lv, uint32_t v, uint32_t sp, uint32_t ep, bool active,
                                           int h, uint32_t px, uint32_t& rs, uint32_t& re, const IndexView& iv) {
  constexpr uint32_t kSyms = Line3<kM>::kSyms;
  const bool special = v == kSpecialCode;
  const uint32_t ls = dna_line_of_t<kM>(sp), le = dna_line_of_t<kM>(ep);
  const uint32_t os = sp - ls * kSyms, oe = ep - le * kSyms;
  const bool load = active && !special;
  const bool split = load && (le != ls);
  Chunk32 ks = chunk_undefined(), ke = chunk_undefined();
  if (load) { check_line(iv, lv + (size_t)ls * kLine3Bytes, 32); ks = ldg_line_keep(lv + (size_t)ls * kLine3Bytes); }
  if (split) { check_line(iv, lv + (size_t)le * kLine3Bytes, 32); ke = ldg_line_keep(lv + (size_t)le * kLine3Bytes); }
  DnaHits xs = dna_hits<kM>(ks, v);
  const uint32_t cs = dna_counter<kM>(ks, v, h, ls * kSyms);
  const uint32_t ps = dna_partial<kM>(cs, xs, os, h);
  uint32_t ce = cs;
  if (__any_sync(0xFFFFFFFFu, split)) {  // warp-uniform: skipped once every interval is narrower than a line
    const DnaHits x2 = dna_hits<kM>(ke, v);
    const uint32_t c2 = dna_counter<kM>(ke, v, h, le * kSyms);
    xs.h0 = split ? x2.h0 : xs.h0;
    xs.h1 = split ? x2.h1 : xs.h1;
    xs.h2 = split ? x2.h2 : xs.h2;
    ce = split ? c2 : cs;
  }
  const uint32_t pe = dna_partial<kM>(ce, xs, oe, h);
  rs = group2_sum(ps);
  re = group2_sum(pe);
  // the symbol that occurs once is stored (and counted) as a 0: rank(0, p) is one too high beyond its row
  const uint32_t zero = v == 0u ? 1u : 0u;
  rs -= zero & (sp > px ? 1u : 0u);
  re -= zero & (ep > px ? 1u : 0u);
  if (special) {
    rs = sp > px ? 1u : 0u;
    re = ep > px ? 1u : 0u;
  }
}

// The same for ONE lane per query (count3_kernel<., 1>, indexes that live in the L2): the lane loads the whole 64-byte
// line itself (two 256-bit loads) and needs no shuffle; a warp carries 32 queries, so everything around the rank (loop
// control, refill, step set-up, address arithmetic) costs half the warp instructions per query. Out of HBM this form
// would spend two fetch slots per line (profiles/README.md §R2.3), so it is used only while the index fits the L2.
template <bool kM>
__device__ __forceinline__ void rank_pair3t(const uint8_t* __restrict__ lv, uint32_t v, uint32_t sp, uint32_t ep, bool active,
                                            uint32_t px, uint32_t& rs, uint32_t& re, const IndexView& iv) {
  constexpr uint32_t kSyms = Line3<kM>::kSyms;
  const bool special = v == kSpecialCode;
  const uint32_t ls = dna_line_of_t<kM>(sp), le = dna_line_of_t<kM>(ep);
  const uint32_t os = sp - ls * kSyms, oe = ep - le * kSyms;
  const bool load = active && !special;
  const bool split = load && (le != ls);
  Chunk32 A = chunk_undefined(), B = chunk_undefined();
  if (load) {
    check_line(iv, lv + (size_t)ls * kLine3Bytes, 64);
    A = ldg_line_keep(lv + (size_t)ls * kLine3Bytes);
    B = ldg_line_keep(lv + (size_t)ls * kLine3Bytes + 32);
  }
  uint32_t x[6], cnt;
  dna_line_hits<kM>(A, B, v, x, cnt, ls * kSyms);
  rs = dna_line_rank<kM>(cnt, x, os);
  re = dna_line_rank<kM>(cnt, x, oe);
  if (__any_sync(0xFFFFFFFFu, split)) {  // warp-uniform: rare once the intervals are narrower than a line
    if (split) {
      check_line(iv, lv + (size_t)le * kLine3Bytes, 64);
      A = ldg_line_keep(lv + (size_t)le * kLine3Bytes);
      B = ldg_line_keep(lv + (size_t)le * kLine3Bytes + 32);
    }
    uint32_t y[6], c2;
    dna_line_hits<kM>(A, B, v, y, c2, le * kSyms);
    const uint32_t r2 = dna_line_rank<kM>(c2, y, oe);
    re = split ? r2 : re;
  }
  const uint32_t zero = v == 0u ? 1u : 0u;  // the symbol that occurs once is stored (and counted) as a 0
  rs -= zero & (sp > px ? 1u : 0u);
  re -= zero & (ep > px ? 1u : 0u);
  if (special) {
    rs = sp > px ? 1u : 0u;
    re = ep > px ? 1u : 0u;
  }
}

// ------------------------------------------------------------------------------------------
// count (fm_index.cpp:79-101)
// ------------------------------------------------------------------------------------------
#ifndef CSFM_COUNT3_CTAS
#define CSFM_COUNT3_CTAS 6  // measured on C2: 6 (40 registers, no spill) beats 8 by 4 %
#endif
#ifndef CSFM_COUNT3T_CTAS
#define CSFM_COUNT3T_CTAS 4  // measured on C2: 1.076 / 1.070 / 1.081 / 0.651e10 q/s at 3 / 4 / 5 / 6 (6 spills)
#endif
// kLanes = 2: a two-lane sub-warp per query (16 per warp, one 64-byte request per line: any index size);
// kLanes = 1: one lane per query (32 per warp: indexes that fit the L2).
template <bool kInstr, int kLanes, bool kM>
__global__ void __launch_bounds__(kThreads, kLanes == 2 ? CSFM_COUNT3_CTAS : CSFM_COUNT3T_CTAS)
count3_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ CountArgs a) {
  __shared__ uint32_t sC[257];
  __shared__ uint2 step_tab[256];  // x = C[byte], y = compact code | (byte absent) << 31
  for (int i = threadIdx.x; i < 257; i += blockDim.x) sC[i] = iv.hdr->C[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    step_tab[i] = make_uint2(iv.hdr->C[i], iv.hdr->code_of_byte[i] | (iv.hdr->C[i + 1] == iv.hdr->C[i] ? 0x80000000u : 0u));
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int h = kLanes == 2 ? (lane & 1) : 0;
  const bool leader = h == 0;  // the lane that writes the query's results
  const uint8_t* const lv = iv.levels + h * 32;
  const uint32_t px = iv.special_row;
  const uint32_t kk = iv.kmer_k;
  WarpQueue32 wq;

  bool active = false;
  uint32_t q = 0;
  const uint8_t* ptr = nullptr;  // address of the character being processed
  uint32_t rem = 0;              // characters left including the current one
  uint32_t sp = 0, ep = 0, base = 0, code = 0, next_byte = 0;
  uint32_t my_steps = 0, my_lookups = 0, my_lines = 0;

  auto finish = [&](uint32_t cnt, uint32_t lo, uint32_t hi) {
    if (leader) {
      if (a.counts) a.counts[q] = cnt;
      if (a.sp_ep) {
        a.sp_ep[2 * (uint64_t)q] = lo;
        a.sp_ep[2 * (uint64_t)q + 1] = hi;
      }
      if (a.row_sp) {
        a.row_sp[q] = lo;
        a.row_cnt[q] = cnt < a.limit32 ? cnt : a.limit32;
      }
    }
    active = false;
  };
  auto begin_step = [&](uint32_t b) {  // sets up the step that prepends byte b to [sp, ep)
    if (kInstr) ++my_steps;
    const uint2 e = step_tab[b];
    if (e.y & 0x80000000u) {  // symbol absent: occ(c,.) == 0 -> sp == ep (fm_index.cpp:96)
      finish(0, 0, 0);
      return;
    }
    code = e.y;
    base = e.x;
    if (rem > 1) next_byte = ptr[-1];  // prefetch: in flight during the rank
  };

  for (;;) {
    const uint32_t item = queue_take32g<kLanes>(wq, !active, lane, a.cursor, (uint32_t)a.npat);
    if (item != ~0u) {
      q = item;
      const uint64_t o0 = a.offs[q], o1 = a.offs[(uint64_t)q + 1];
      const uint64_t m = o1 - o0;
      CSFM_CHK(q < a.npat && o0 <= o1 && o1 <= a.offs[a.npat], "pattern inside the batch");
      active = true;
      if (m == 0) {
        // count("") == n (fm_index.cpp:80); locate("") is empty (fm_index.cpp:109)
        if (leader) {
          if (a.counts) a.counts[q] = iv.n;
          if (a.sp_ep) { a.sp_ep[2 * (uint64_t)q] = 0; a.sp_ep[2 * (uint64_t)q + 1] = 0; }
          if (a.row_sp) { a.row_sp[q] = 0; a.row_cnt[q] = 0; }
        }
        active = false;
      } else {
        // k-mer jump table: the interval of the last k characters in one lookup. Keys are made of the two-bit
        // codes; a pattern whose last k characters contain the symbol that occurs once takes the first step
        // from the C array like a short pattern.
        uint32_t e = 0, mul = 1;
        bool present = true, keyed = kk != 0 && m >= kk;
        if (keyed) {
          for (uint32_t i = 0; i < kk; ++i) {
            const uint2 t = step_tab[a.bytes[o1 - 1 - i]];
            present = present && !(t.y & 0x80000000u);
            keyed = keyed && (t.y & 0xFFu) != kSpecialCode;
            e += (t.y & 3u) * mul;
            mul *= iv.kmer_radix;
          }
        }
        if (!present) {
          if (kInstr && keyed) ++my_lookups;  // answered without a step: one of the k characters does not occur
          finish(0, 0, 0);
        } else if (keyed) {
          if (kInstr) ++my_lookups;
          CSFM_CHK(e < iv.kmer_entries, "k-mer key inside the table");
          const uint2 se = ldg_table_keep(&iv.kmer[e]);
          sp = se.x;
          ep = se.y;
          if (sp >= ep) {
            finish(0, 0, 0);
          } else if (m == kk) {
            finish(ep - sp, sp, ep);
          } else {
            rem = (uint32_t)(m - kk);
            ptr = a.bytes + (o1 - kk - 1);
            begin_step(*ptr);
          }
        } else {
          // first step needs no rank: occ(c,0) = 0 and occ(c,n) = freq[c]  => [C[c], C[c+1])
          const uint32_t b = a.bytes[o1 - 1];
          sp = sC[b];
          ep = sC[b + 1];
          if (kInstr) ++my_steps;
          if (sp >= ep) {
            finish(0, 0, 0);
          } else if (m == 1) {
            finish(ep - sp, sp, ep);
          } else {
            rem = (uint32_t)(m - 1);
            ptr = a.bytes + (o1 - 2);
            begin_step(*ptr);
          }
        }
      }
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- one backward-search step: sp/ep <- C[c] + rank(c, .)  (fm_index.cpp:92-93)
    if (kInstr && active && code != kSpecialCode) my_lines += 1u + (dna_line_of_t<kM>(sp) != dna_line_of_t<kM>(ep) ? 1u : 0u);
    uint32_t rs, re;
    if constexpr (kLanes == 2) rank_pair3<kM>(lv, code, sp, ep, active, h, px, rs, re, iv);
    else rank_pair3t<kM>(lv, code, sp, ep, active, px, rs, re, iv);
    if (active) {
      sp = base + rs;
      ep = base + re;
      if (sp >= ep) {
        finish(0, 0, 0);
      } else if (--rem == 0) {
        finish(ep - sp, sp, ep);
      } else {
        --ptr;
        begin_step(next_byte);
      }
    }
  }
  if (kInstr && a.steps_total) {
    unsigned s = leader ? my_steps : 0, t = leader ? my_lookups : 0, w = leader ? my_lines : 0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
      w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
    }
    if (lane == 0 && s) atomicAdd(a.steps_total, (unsigned long long)s);
    if (lane == 0 && t) atomicAdd(a.steps_total + 1, (unsigned long long)t);
    if (lane == 0 && w) atomicAdd(a.steps_total + 4, (unsigned long long)w);
  }
}

// Fills the k-mer jump table by running the same backward search over every k-symbol string:
// entry e decodes to codes d_0 (LAST character, e % radix), d_1, ... and stores its interval.
template <bool kM>
__global__ void __launch_bounds__(kThreads)
kmer_build3_kernel(const __grid_constant__ IndexView iv, uint2* __restrict__ table, unsigned long long entries, uint32_t k,
                   uint32_t radix) {
  __shared__ uint32_t sC[257];
  __shared__ uint8_t byte_of_code[8];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) sC[i] = iv.hdr->C[i];
  if (threadIdx.x < 8) byte_of_code[threadIdx.x] = iv.hdr->byte_of_code[threadIdx.x];
  __syncthreads();
  const int h = threadIdx.x & 1;
  const uint8_t* const lv = iv.levels + h * 32;
  const uint32_t px = iv.special_row;
  const unsigned long long group = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const unsigned long long ngroups = ((unsigned long long)gridDim.x * blockDim.x) >> 1;
  const unsigned long long trips = (entries + ngroups - 1) / ngroups;
  for (unsigned long long t = 0; t < trips; ++t) {
    const unsigned long long e = t * ngroups + group;
    const bool valid = e < entries;
    unsigned long long rest = valid ? e : 0;
    uint32_t byte = byte_of_code[rest % radix];
    rest /= radix;
    uint32_t sp = sC[byte], ep = sC[byte + 1];  // first step: [C[c], C[c+1])
    bool alive = valid && sp < ep;
    for (uint32_t i = 1; i < k; ++i) {
      const uint32_t code = (uint32_t)(rest % radix);
      rest /= radix;
      byte = byte_of_code[code];
      const bool act = alive && (sC[byte + 1] != sC[byte]);
      uint32_t rs, re;
      rank_pair3<kM>(lv, code, sp, ep, act, h, px, rs, re, iv);
      sp = sC[byte] + rs;
      ep = sC[byte] + re;
      alive = act && sp < ep;
    }
    if (valid && h == 0) table[e] = alive ? make_uint2(sp, ep) : make_uint2(0u, 0u);
  }
}

// One LF step's memory access: the symbol at row p and its rank before p, from ONE line. Marked form: also the mark bit
// of row p and the number of marked rows before it (the index of its position sample).
template <bool kM>
__device__ __forceinline__ uint32_t access_rank3(const uint8_t* __restrict__ lv, uint32_t p, bool active, int lane, int h,
                                                 uint32_t& v, const IndexView& iv, uint32_t* mark = nullptr, uint32_t* mrank = nullptr) {
  constexpr uint32_t kSyms = Line3<kM>::kSyms;
  Chunk32 k = chunk_undefined();
  const uint32_t line = dna_line_of_t<kM>(p), off = p - line * kSyms;
  if (active) { check_line(iv, lv + (size_t)line * kLine3Bytes, 32); k = ldg_line_keep(lv + (size_t)line * kLine3Bytes); }
  const int owner = (lane & ~1) | (off >= Line3<kM>::kHalf ? 1 : 0);
  if constexpr (kM) {
    // symbol and mark bit travel in one shuffle
    const uint32_t got = __shfl_sync(0xFFFFFFFFu, dna_symbol<true>(k, off, h) | (dna_mark(k, off, h) << 2), owner);
    v = got & 3u;
    if (mark) {
      *mark = got >> 2;
      *mrank = group2_sum(dna_mark_partial(k, off, h));
    }
  } else {
    v = __shfl_sync(0xFFFFFFFFu, dna_symbol<false>(k, off, h), owner);
  }
  return group2_sum(dna_partial<kM>(dna_counter<kM>(k, v, h, line * kSyms), dna_hits<kM>(k, v), off, h));
}

// The same for one LANE per row (walk3_kernel<1, .>): the lane loads both halves of the line, no shuffles.
template <bool kM>
__device__ __forceinline__ uint32_t access_rank3t(const uint8_t* __restrict__ lv, uint32_t p, bool active, uint32_t& v,
                                                  const IndexView& iv, uint32_t* mark = nullptr, uint32_t* mrank = nullptr) {
  constexpr uint32_t kSyms = Line3<kM>::kSyms;
  Chunk32 A = chunk_undefined(), B = chunk_undefined();
  const uint32_t line = dna_line_of_t<kM>(p), off = p - line * kSyms;
  if (active) {
    check_line(iv, lv + (size_t)line * kLine3Bytes, 64);
    A = ldg_line_keep(lv + (size_t)line * kLine3Bytes);
    B = ldg_line_keep(lv + (size_t)line * kLine3Bytes + 32);
  }
  v = dna_line_symbol<kM>(A, B, off);
  if constexpr (kM) {
    if (mark) {
      *mark = dna_line_mark(A, B, off);
      *mrank = dna_line_mark_rank(A, B, off);
    }
  }
  uint32_t x[6], cnt;
  dna_line_hits<kM>(A, B, v, x, cnt, line * kSyms);
  return dna_line_rank<kM>(cnt, x, off);
}

// ------------------------------------------------------------------------------------------
// locate: rows -> text positions (fm_index.cpp:125-153, LF of fm_index.hpp:62-66)
// ------------------------------------------------------------------------------------------
#ifndef CSFM_WALK3_CTAS
#define CSFM_WALK3_CTAS 6
#endif
#ifndef CSFM_WALK3T_CTAS
#define CSFM_WALK3T_CTAS 6  // measured on C4: 3.37 / 3.33 / 3.28 / 3.48e9 occ/s at 3 / 4 / 5 / 6 (39 registers; 7 spills)
#endif
// kLanes = 2: a two-lane sub-warp per row (one 64-byte request per line); kLanes = 1: one lane per row (32 walks per warp)
// kM: the marked line form — the walk ends at the first row whose suffix starts at a multiple of the stride (its mark bit
// sits in the line the step fetches anyway): at most stride - 1 steps, SA[row] = sample + steps as before.
template <int kLanes, bool kM>
__global__ void __launch_bounds__(kThreads, kLanes == 2 ? CSFM_WALK3_CTAS : CSFM_WALK3T_CTAS)
walk3_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ WalkArgs a) {
  __shared__ uint32_t base_by_code[8];
  if (threadIdx.x < 8) base_by_code[threadIdx.x] = iv.hdr->base_by_code[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int h = kLanes == 2 ? (lane & 1) : 0;
  const uint8_t* const lv = iv.levels + h * 32;
  const uint32_t px = iv.special_row;
  WarpQueue wq;

  bool active = false;
  unsigned long long slot = 0;
  uint32_t start = 0, p = 0, steps = 0;
  uint32_t my_lf = 0;

  // Where the reference throws, the whole query fails: attribute the slot to its query.
  auto fail_walk = [&](int why) {
    if (h == 0) {
      unsigned long long lo = 0, hi_q = a.npat;  // last q with out_offs[q] <= slot
      while (hi_q - lo > 1) {
        const unsigned long long mid = (lo + hi_q) >> 1;
        if (a.out_offs[mid] <= slot) lo = mid; else hi_q = mid;
      }
      if (a.status) atomicMax(&a.status[lo], why);
      a.out_pos[slot] = 0;
    }
    active = false;
  };
  auto emit = [&](uint32_t row) {  // row is sampled: SA[row] = ssa[row/stride]; marked form: row = index of the position sample
    const uint32_t k = kM ? row : sample_index(iv, row);
    if (k >= iv.nsamp) {  // fm_index.cpp:141-146 (unreachable for a consistent index)
      fail_walk((int)CSFM_Q_SSA_OOB);
      return;
    }
    CSFM_CHK(k < iv.nsamp, "sample index inside the sampled suffix array");
    if (h == 0) {
      uint64_t pos = (uint64_t)ldg_once_u32(kM ? &iv.psamp[k] : &iv.ssa[k]) + steps;  // fm_index.cpp:147-152
      if (pos >= iv.n) pos -= iv.n;                // sa_val < n and steps < n
      stg_once_u64(&a.out_pos[slot], pos);
    }
    active = false;
  };

  for (;;) {
    const unsigned long long item = queue_takeg<kLanes>(wq, !active, lane, a.cursor, a.total);
    if (item != ~0ull) {
      slot = a.first + item;
      start = a.rows_implicit ? a.row_base + (uint32_t)slot : (uint32_t)a.out_pos[slot];
      steps = 0;
      active = true;
      p = start;
      if constexpr (!kM)
        if (row_is_sampled(iv, start)) emit(start);
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- one LF step: LF(i) = C[c] + occ(c, i)  (fm_index.hpp:62-66)
    uint32_t v, mk = 0, mr = 0;
    uint32_t r;
    if constexpr (kLanes == 2) r = access_rank3<kM>(lv, p, active, lane, h, v, iv, &mk, &mr);
    else r = access_rank3t<kM>(lv, p, active, v, iv, &mk, &mr);
    if (kM && active && mk) {
      emit(mr);  // row p is marked: SA[p] = psamp[mr], the walk started `steps` positions further on
    } else if (active) {
      // row px holds the symbol that occurs once: LF(px) = C[that symbol] + 0
      const uint32_t row = p == px ? iv.special_first : base_by_code[v] + r - ((v == 0u && p > px) ? 1u : 0u);
      ++steps;
      ++my_lf;
      if (kM) {
        // a consistent marked index ends every walk within stride - 1 steps; a corrupt one must not spin
        if (steps >= iv.n) fail_walk((int)CSFM_Q_LF_WALK_EXCEEDED);
        else p = row;
      } else if (row_is_sampled(iv, row)) {
        emit(row);
      } else if (row == start || steps >= iv.n) {
        // LF is a permutation: back at the start without meeting a sampled row means the
        // reference would walk n steps and throw (fm_index.cpp:130-138).
        fail_walk((int)CSFM_Q_LF_WALK_EXCEEDED);
      } else {
        p = row;
      }
    }
  }
  if (a.lf_total) {
    unsigned s = (h == 0) ? my_lf : 0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0 && s) atomicAdd(a.lf_total, (unsigned long long)s);
  }
}

// untext (see untext2_kernel in csfm_query2.cu): the text back out of the index, one two-lane sub-warp per sampled row
template <bool kM>
__global__ void __launch_bounds__(kThreads, 8)
untext3_kernel(const __grid_constant__ IndexView iv, uint8_t* __restrict__ out, unsigned long long* __restrict__ cursor,
               unsigned long long* __restrict__ written) {
  __shared__ uint32_t base_by_code[8];
  __shared__ uint8_t byte_of_code[8];
  if (threadIdx.x < 8) {
    base_by_code[threadIdx.x] = iv.hdr->base_by_code[threadIdx.x];
    byte_of_code[threadIdx.x] = iv.hdr->byte_of_code[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int h = lane & 1;
  const uint8_t* const lv = iv.levels + h * 32;
  const uint32_t px = iv.special_row;
  WarpQueue wq;
  bool active = false;
  uint32_t p = 0, pos = 0, steps = 0, my_written = 0;
  for (;;) {
    const unsigned long long item = queue_takeg<2>(wq, !active, lane, cursor, iv.nsamp);
    if (item != ~0ull) {
      p = (uint32_t)item * iv.stride;
      pos = iv.ssa[item];
      steps = 0;
      active = true;
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;
    uint32_t v;
    const uint32_t r = access_rank3<kM>(lv, p, active, lane, h, v, iv);
    if (active) {
      pos = pos == 0 ? iv.n - 1 : pos - 1;  // BWT[row] = T[(SA[row] - 1) mod n]  (bwt.hpp:10-13)
      if (h == 0) out[pos] = p == px ? (uint8_t)iv.special_byte : byte_of_code[v];
      ++my_written;
      p = p == px ? iv.special_first : base_by_code[v] + r - ((v == 0u && p > px) ? 1u : 0u);
      if (row_is_sampled(iv, p) || ++steps >= iv.n) active = false;
    }
  }
  unsigned s = (h == 0) ? my_written : 0;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0 && s) atomicAdd(written, (unsigned long long)s);
}

// ------------------------------------------------------------------------------------------
// access: BWT[i] for all i (wavelet.cpp:102-128) — verification / export, not a query path
// ------------------------------------------------------------------------------------------
template <bool kM>
__global__ void __launch_bounds__(kThreads)
access3_kernel(const __grid_constant__ IndexView iv, uint8_t* __restrict__ out) {
  __shared__ uint8_t byte_of_code[8];
  if (threadIdx.x < 8) byte_of_code[threadIdx.x] = iv.hdr->byte_of_code[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int h = lane & 1;
  const uint8_t* const lv = iv.levels + h * 32;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 1;
  const uint64_t trips = ((uint64_t)iv.n + ngroups - 1) / ngroups;
  for (uint64_t t = 0; t < trips; ++t) {
    const uint64_t i = t * ngroups + group;
    const bool valid = i < iv.n;
    uint32_t v;
    (void)access_rank3<kM>(lv, valid ? (uint32_t)i : 0u, valid, lane, h, v, iv);
    if (valid && h == 0) out[i] = (uint32_t)i == iv.special_row ? (uint8_t)iv.special_byte : byte_of_code[v];
  }
}

int blocks_per_sm3(const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm;
}

}  // namespace

namespace {
template <bool kM>
const void* count3_fn(bool instr, int lanes) {
  if (lanes == 1) return instr ? (const void*)count3_kernel<true, 1, kM> : (const void*)count3_kernel<false, 1, kM>;
  return instr ? (const void*)count3_kernel<true, 2, kM> : (const void*)count3_kernel<false, 2, kM>;
}
template <bool kM>
void launch_count3_t(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream, int lanes) {
  if (lanes == 1) {
    if (a.steps_total) count3_kernel<true, 1, kM><<<grid, kThreads, 0, stream>>>(iv, a);
    else count3_kernel<false, 1, kM><<<grid, kThreads, 0, stream>>>(iv, a);
  } else {
    if (a.steps_total) count3_kernel<true, 2, kM><<<grid, kThreads, 0, stream>>>(iv, a);
    else count3_kernel<false, 2, kM><<<grid, kThreads, 0, stream>>>(iv, a);
  }
}
template <bool kM>
const void* walk3_fn(int lanes) { return lanes == 1 ? (const void*)walk3_kernel<1, kM> : (const void*)walk3_kernel<2, kM>; }
}  // namespace

void launch_count3(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream, int lanes) {
  if (iv.marked) launch_count3_t<true>(iv, a, grid, stream, lanes);
  else launch_count3_t<false>(iv, a, grid, stream, lanes);
}
void launch_walk3(const IndexView& iv, const WalkArgs& a, int grid, cudaStream_t stream, int lanes) {
  if (iv.marked) {
    if (lanes == 1) walk3_kernel<1, true><<<grid, kThreads, 0, stream>>>(iv, a);
    else walk3_kernel<2, true><<<grid, kThreads, 0, stream>>>(iv, a);
  } else {
    if (lanes == 1) walk3_kernel<1, false><<<grid, kThreads, 0, stream>>>(iv, a);
    else walk3_kernel<2, false><<<grid, kThreads, 0, stream>>>(iv, a);
  }
}
void launch_untext3(const IndexView& iv, uint8_t* out, unsigned long long* cursor, unsigned long long* written, int num_sms,
                    cudaStream_t stream) {
  const unsigned long long want = ((unsigned long long)iv.nsamp * 2 + kThreads - 1) / kThreads;
  const void* fn = iv.marked ? (const void*)untext3_kernel<true> : (const void*)untext3_kernel<false>;
  const int grid = (int)std::min<unsigned long long>(std::max<unsigned long long>(want, 1), (unsigned long long)num_sms * blocks_per_sm3(fn));
  if (iv.marked) untext3_kernel<true><<<grid, kThreads, 0, stream>>>(iv, out, cursor, written);
  else untext3_kernel<false><<<grid, kThreads, 0, stream>>>(iv, out, cursor, written);
}
void launch_access3(const IndexView& iv, uint8_t* out, int grid, cudaStream_t stream) {
  if (iv.marked) access3_kernel<true><<<grid, kThreads, 0, stream>>>(iv, out);
  else access3_kernel<false><<<grid, kThreads, 0, stream>>>(iv, out);
}
int max_blocks_per_sm_count3(const IndexView& iv, const CountArgs& a, int lanes) {
  return blocks_per_sm3(iv.marked ? count3_fn<true>(a.steps_total != nullptr, lanes) : count3_fn<false>(a.steps_total != nullptr, lanes));
}
int max_blocks_per_sm_walk3(const IndexView& iv, int lanes) { return blocks_per_sm3(iv.marked ? walk3_fn<true>(lanes) : walk3_fn<false>(lanes)); }
int max_blocks_per_sm_access3(const IndexView& iv) {
  return blocks_per_sm3(iv.marked ? (const void*)access3_kernel<true> : (const void*)access3_kernel<false>);
}

int build_kmer_table3(csfm_index* idx, cudaStream_t stream) {
  const BlobHeader& h = idx->h;
  if (!h.kmer_k) return CSFM_OK;
  unsigned long long entries = 1;
  for (uint32_t i = 0; i < h.kmer_k; ++i) entries *= h.kmer_radix;
  uint2* table = reinterpret_cast<uint2*>(idx->d_blob + h.off_kmer);
  const unsigned long long want = (entries * 2 + kThreads - 1) / kThreads;
  IndexView v = idx->view;  // the table is being written: the builder itself must not consult it
  v.kmer = nullptr;
  v.kmer_k = 0;
  const void* fn = v.marked ? (const void*)kmer_build3_kernel<true> : (const void*)kmer_build3_kernel<false>;
  const int grid = (int)std::min<unsigned long long>(want, (unsigned long long)idx->num_sms * blocks_per_sm3(fn));
  if (v.marked) kmer_build3_kernel<true><<<grid, kThreads, 0, stream>>>(v, table, entries, h.kmer_k, h.kmer_radix);
  else kmer_build3_kernel<false><<<grid, kThreads, 0, stream>>>(v, table, entries, h.kmer_k, h.kmer_radix);
  CSFM_CUDA(cudaGetLastError());
  CSFM_CUDA(cudaStreamSynchronize(stream));
  return CSFM_OK;
}

}  // namespace csfm
