// csfm_query2.cu — query kernels for layout 2 (16-ary levels in 128-byte lines), sm_100a.
//
// Execution model: a 4-lane sub-warp owns one query (count) / one occurrence row (locate). One
// rank = one 128-byte line = four 256-bit loads (LDG.E.256, one per lane, ONE L1 wavefront), a
// counter pick, one hit word (four logic ops on the bit-sliced payload) + one masked popc per lane,
// two xor-shuffles. A byte
// alphabet needs two dependent line fetches per rank, DNA one.
//
// All eight sub-warps of a warp run one loop in lock-step and in PHASE: a loop trip is a whole
// backward-search step (level 0, level 1, step end), so the step-end code (interval test, next
// pattern byte, table lookups) is issued once per step for the warp, not once per level per
// straggler. A sub-warp that finishes refills at the next trip from a warp-local chunk of the
// batch cursor (ballot-ranked, one global atomic per 32 queries). sp and ep usually fall in the
// same line after ~4 steps: then one load and one hit word serve both ends, and the warp skips
// the second hit computation altogether when no sub-warp needs it. Grids are persistent.
//
// A query does not start at its last character: the k-mer jump table gives the interval of its last
// k characters in one lookup; on two-level indexes the half-step table gives that interval already
// mapped through level 0 for the next character's high nibble, so the first rank step reads only its
// level-1 line. When the index carries the text and the suffix array, a query whose interval is down
// to at most four rows finishes by comparing its remaining characters with the text in front of those
// suffixes (count2_kernel<true,.>); with a large table the lookup itself may leave so few rows.
//
// Replaces cs::FMIndex::count / locate (/root/reference/src/api/fm_index.cpp:79-157),
// cs::WaveletTree::rank / access (src/core/wavelet.cpp:59-128) and cs::BitVector::rank1
// (src/core/bitvector.cpp:165-230).
#include <algorithm>

#include <cub/device/device_scan.cuh>

#include "csfm_host.hpp"
#include "csfm_kernels.cuh"

namespace csfm {

namespace {

// rank_l(v, sp) and rank_l(v, ep) for one level. lv = level base + 32*j. All 32 lanes call it.
__device__ __forceinline__ void rank_pair(const uint8_t* __restrict__ lv, uint32_t v, uint32_t sp_pos, uint32_t ep_pos,
                                          bool active, int j, uint32_t& rs, uint32_t& re, const IndexView& iv) {
  const uint32_t ls = sp_pos & ~(kSymsPerLine - 1), le = ep_pos & ~(kSymsPerLine - 1);  // line*128 == byte offset
  const uint32_t os = sp_pos - ls, oe = ep_pos - le;
  const bool split = active && (le != ls);
  Chunk32 ks = chunk_undefined(), ke = chunk_undefined();
  if (active) { check_line(iv, lv + ls, 32); ks = ldg_nc_v8(lv + ls); }
  if (split) { check_line(iv, lv + le, 32); ke = ldg_nc_v8(lv + le); }
  const uint32_t hs = chunk_hits(ks, v);
  const uint32_t cs = chunk_counter(ks, v, j);
  uint32_t he = hs, ce = cs;
  if (__any_sync(0xFFFFFFFFu, split)) {  // warp-uniform: skipped once every interval is narrow
    uint32_t v2 = v;
    asm("" : "+r"(v2));  // opaque copy: rebuild the four plane masks here instead of keeping them in registers
    const uint32_t h2 = chunk_hits(ke, v2);
    const uint32_t c2 = chunk_counter(ke, v, j);
    he = split ? h2 : hs;
    ce = split ? c2 : cs;
  }
  rs = group4_sum(chunk_partial(cs, hs, os, j));
  re = group4_sum(chunk_partial(ce, he, oe, j));
}

// ------------------------------------------------------------------------------------------
// count (fm_index.cpp:79-101)
// ------------------------------------------------------------------------------------------
// Default variant: pattern offsets and bytes are read straight from global memory (two dependent
// loads when a sub-warp starts a pattern, then a one-step byte prefetch).
// kShortcut: compiled with the text-verification stages (used when the index carries the text, no
// interval is asked for and the batch is 16-byte aligned); the plain variant keeps 32 registers.
#ifndef CSFM_SHORTCUT_CTAS
#define CSFM_SHORTCUT_CTAS 6  // measured: 6 beats 5 (no spills) and 8 (more spills)
#endif
// kInstr: compiled with the per-call step / lookup / verification counters (csfm_set_instrumentation).
template <bool kShortcut, bool kInstr>
__global__ void __launch_bounds__(kThreads, kShortcut ? CSFM_SHORTCUT_CTAS : 8)
count2_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ CountArgs a) {
  __shared__ Tables tb;
  __shared__ uint4 step_tab[256];
  if (a.qlist && *a.qlist_len == 0) return;  // second pass with nothing left (uniform for the whole grid)
  load_step_table(step_tab, iv.hdr);
  load_tables(tb, iv.hdr);

  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv_last = iv.levels_last + j * 32;
  const uint32_t kk = iv.kmer_k;  // 0 = no jump table
  WarpQueue32 wq;
  // second pass of a two-pass count: the items are the queries the first pass left in a.qlist
  const uint32_t nitems = a.qlist ? (uint32_t)*a.qlist_len : (uint32_t)a.npat;

  // Text-verification shortcut (see csfm_common.cuh): only when no interval is asked for (a
  // verified query yields its count, 0 or 1, but not its final SA row) and the batch bytes are
  // 16-byte aligned (the pattern window is staged with 16-byte cp.async copies).
  __shared__ VerifySlot vslots[kShortcut ? kThreads / 4 : 1];
  VerifySlot& vs = vslots[kShortcut ? (threadIdx.x >> 2) : 0];
  constexpr bool shortcut = kShortcut;
  const uint8_t* const batch_end = shortcut ? a.bytes + a.offs[a.npat] : nullptr;
  const uint32_t dense_mask = (1u << iv.dense_shift) - 1u;
  const uint32_t max_rows = iv.dense_shift == 0 ? kVerifyRows : 1u;  // neighbouring rows need the full suffix array

  bool active = false;
  uint32_t q = 0;                // query index (a launch stays below 2^32 queries)
  const uint8_t* ptr = nullptr;  // address of the character being processed
  uint32_t rem = 0;              // characters left including the current one
  uint32_t sp = 0, ep = 0, base = 0, add0 = 0, code = 0, next_byte = 0;
  uint32_t vstage = 0, vp = 0;   // 0 searching, 1 suffix-array entry requested, 2 windows in flight
  uint32_t my_steps = 0, my_lookups = 0, my_checks = 0, my_halves = 0, my_lines = 0;
  uint32_t since_refill = 0;     // warp-uniform: trips since this warp last took queries
  bool pre = false;              // sp/ep already hold the level-1 interval of the pending step (half-step table)
  const uint2* const kmer_hi = kShortcut ? iv.kmer_hi : nullptr;  // the plain variant stays within 32 registers

  auto finish = [&](uint32_t cnt, uint32_t lo, uint32_t hi) {
    if (j == 0) {
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
  // Sets up the step that prepends byte b to the interval [sp,ep).
  auto begin_step = [&](uint32_t b) {
    if (kInstr) ++my_steps;
    const uint4 e = step_tab[b];
    if (e.y & 0x80000000u) {  // symbol absent: occ(c,.) == 0 -> sp == ep (fm_index.cpp:96)
      finish(0, 0, 0);
      return;
    }
    code = e.y;
    base = e.x;
    add0 = e.z;
    if (rem > 1) next_byte = ptr[-1];  // prefetch: in flight during the rank levels
  };

  for (;;) {
    // ---- refill (every sub-warp is at a step boundary here) -------------------------------
    // The shortcut variant refills in generations: the queries of a warp then move through the
    // stages together (refill+step, suffix-array entry, windows, compare), so a trip runs only
    // the sections its generation is in. A generation is not held up for more than refill_wait
    // trips by a query that keeps stepping.
    bool served;
    // (the stepping-only variant keeps per-sub-warp refills: batching them was measured to gain
    // nothing there -- equal-length batches fall into step by themselves -- and its bookkeeping
    // cost 12 % on the issue-bound C2 workload)
    const uint32_t item = queue_take32(wq, !active, lane, a.cursor, nitems,
                                       kShortcut && since_refill < iv.refill_wait ? iv.refill_min : 1u, served);
    if (kShortcut) since_refill = served ? 0u : since_refill + 1u;
    if (item != ~0u) {
      q = a.qlist ? a.qlist[item] : item;
      const uint64_t o0 = a.offs[q], o1 = a.offs[(uint64_t)q + 1];
      const uint64_t m = o1 - o0;
      CSFM_CHK(q < a.npat && o0 <= o1 && o1 <= a.offs[a.npat], "pattern inside the batch");
      active = true;
      if (m == 0) {
        // count("") == n (fm_index.cpp:80); locate("") is empty (fm_index.cpp:109)
        if (j == 0) {
          if (a.counts) a.counts[q] = iv.n;
          if (a.sp_ep) { a.sp_ep[2 * (uint64_t)q] = 0; a.sp_ep[2 * (uint64_t)q + 1] = 0; }
          if (a.row_sp) { a.row_sp[q] = 0; a.row_cnt[q] = 0; }
        }
        active = false;
      } else if (kk != 0 && m >= kk) {
        // k-mer jump table: the interval of the last k characters in one lookup (the table was
        // filled by this same backward search, kmer_build_kernel)
        uint32_t e = 0, mul = 1;
        bool present = true;
        for (uint32_t i = 0; i < kk; ++i) {
          const uint32_t b = a.bytes[o1 - 1 - i];
          present = present && (tb.C[b + 1] != tb.C[b]);
          e += tb.code_of_byte[b] * mul;
          mul *= iv.kmer_radix;
        }
        if (kInstr) ++my_lookups;
        uint2 se = make_uint2(0, 0);
        if (kmer_hi != nullptr && m > kk) {
          // half-step table: the k-mer's interval already mapped through level 0 for the high nibble
          // of the character in front of it; the pending step needs only its level-1 rank
          ptr = a.bytes + (o1 - 1 - kk);
          rem = (uint32_t)(m - kk);
          const uint4 st = step_tab[*ptr];
          CSFM_CHK(!present || e < iv.kmer_entries, "k-mer key inside the table");
          if (present && !(st.y & 0x80000000u)) se = kmer_hi[(size_t)e * 16u + (st.y >> 4)];
          sp = se.x;
          ep = se.y;
          if (sp >= ep) {  // the k-mer, the character, or that nibble inside the interval does not occur
            finish(0, 0, 0);
          } else {
            if (kInstr) ++my_halves;
            code = st.y;
            base = st.x;
            if (rem > 1) next_byte = ptr[-1];
            pre = true;
          }
        } else {
          CSFM_CHK(!present || e < iv.kmer_entries, "k-mer key inside the table");
          if (present) {
            if (iv.kmer_tiled) {  // sp only: the next key's sp is this key's ep
              const uint32_t* const t = reinterpret_cast<const uint32_t*>(iv.kmer);
              se = make_uint2(t[e], t[(size_t)e + 1]);
            } else {
              se = iv.kmer[e];
            }
          }
          sp = se.x;
          ep = se.y;
          if (sp >= ep) {
            finish(0, 0, 0);
          } else if (m == kk) {
            finish(ep - sp, sp, ep);
          } else {
            rem = (uint32_t)(m - kk);
            ptr = a.bytes + (o1 - kk);  // the last character the lookup consumed
            if (shortcut && ep - sp <= max_rows && rem >= iv.verify_min && rem <= kVerifyMax && (sp & dense_mask) == 0) {
              // a long table key already leaves few rows: verify them without a rank step
              next_byte = ptr[-1];
              CSFM_CHK(ep <= iv.n, "interval inside the suffix array");
              if ((uint32_t)j < ep - sp) vp = iv.dense[(sp >> iv.dense_shift) + j];
              vstage = 1;
              if (kInstr) ++my_checks;
            } else {
              --ptr;
              begin_step(*ptr);
            }
          }
        }
      } else {
        // first step needs no rank: occ(c,0) = 0 and occ(c,n) = freq[c]  => [C[c], C[c+1])
        const uint32_t b = a.bytes[o1 - 1];
        sp = tb.C[b];
        ep = tb.C[b + 1];
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
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- text verification of narrow intervals, two trips behind the step that found them ----
    // sp and ep stay as that step left them while vstage != 0: lane j owns row sp + j.
    if (kShortcut && __any_sync(0xFFFFFFFFu, vstage != 0)) {  // warp-uniform
      const uint32_t rows = ep - sp;
      if (__any_sync(0xFFFFFFFFu, vstage == 2)) {
        if (vstage == 2) cp_async_wait_all();  // the copies this lane issued last trip
        __syncwarp();                          // ... and those of its three neighbours
        bool hit = false;
        if (vstage == 2 && (uint32_t)j < rows)
          hit = windows_equal(vs.t[j], vs.p, (vp - rem) & 15u, (uint32_t)(reinterpret_cast<uintptr_t>(ptr - rem) & 15u), rem);
        const unsigned votes = __ballot_sync(0xFFFFFFFFu, hit);
        if (vstage == 2) {
          // a row stays in the interval exactly while the characters in front of its suffix match
          vstage = 0;
          finish(__popc((votes >> (lane & ~3)) & 0xFu), 0, 0);
        }
      }
      if (__any_sync(0xFFFFFFFFu, vstage == 1)) {
        const uint8_t* pstart = ptr - rem;
        const uintptr_t pal = reinterpret_cast<uintptr_t>(pstart) & ~(uintptr_t)15;
        const uint32_t np16 = (uint32_t)((reinterpret_cast<uintptr_t>(ptr) - pal + 15) >> 4);
        // an occurrence that would wrap around the text start (cyclic BWT), or an aligned pattern
        // window that would leave the batch: that query goes on stepping
        const bool bad = vstage == 1 && (((uint32_t)j < rows && vp < rem) ||
                                         reinterpret_cast<const uint8_t*>(pal) + 16 * np16 > batch_end);
        const unsigned bads = __ballot_sync(0xFFFFFFFFu, bad);
        const bool stage_ok = vstage == 1 && !((bads >> (lane & ~3)) & 0xFu);
        // text windows: the up to three 16-byte chunks of a row's window go out in ONE instruction, lane c copying
        // chunk c of row r: one memory request per row (rows beyond the first are rare) instead of one per chunk
#pragma unroll
        for (uint32_t r = 0; r < kVerifyRows; ++r) {
          const uint32_t vr = __shfl_sync(0xFFFFFFFFu, vp, (lane & ~3) | (int)r);  // row r's suffix position lives in lane r
          if (r == 0 || __any_sync(0xFFFFFFFFu, stage_ok && r < rows)) {
            const uint32_t tal = (vr - rem) & ~15u;
            const uint32_t nt16 = (vr - tal + 15) >> 4;
            if (stage_ok && r < rows && (uint32_t)j < nt16) {
              CSFM_CHK(vr <= iv.n && vr >= rem && (uint64_t)tal + 16 * j + 16 <= (uint64_t)iv.n + 64, "text window inside the text section");
              cp_async16(&vs.t[r][16 * j], iv.text + tal + 16 * j);
            }
          }
        }
        if (vstage == 1) {
          if ((bads >> (lane & ~3)) & 0xFu) {
            vstage = 0;
            --ptr;
            begin_step(next_byte);
          } else {
            if ((uint32_t)j < np16) cp_async16(&vs.p[16 * j], reinterpret_cast<const uint8_t*>(pal) + 16 * j);
            cp_async_commit();
            vstage = 2;
          }
        }
      }
    }

    // ---- one backward-search step: sp/ep <- base[c] + rank_last(lo, start1[hi] + rank_0(hi, .))
    const bool ranking = active && vstage == 0;
    if (kShortcut && !__any_sync(0xFFFFFFFFu, ranking)) continue;  // a trip of verifications only
    uint32_t rs, re, s1 = sp, e1 = ep;  // sp and ep themselves must survive a trip spent in verification
    if (two && (!kShortcut || __any_sync(0xFFFFFFFFu, ranking && !pre))) {
      if (kInstr && ranking && !pre) my_lines += 1u + (((sp ^ ep) >> 7) != 0u);
      rank_pair(lv0, code >> 4, sp, ep, ranking && !pre, j, rs, re, iv);
      if (!pre) {
        s1 = add0 + rs;
        e1 = add0 + re;
      }
    }
    if (kInstr && ranking) my_lines += 1u + (((s1 ^ e1) >> 7) != 0u);
    rank_pair(lv_last, code & 15u, s1, e1, ranking, j, rs, re, iv);
    if (ranking) {
      pre = false;
      sp = base + rs;  // fm_index.cpp:92-93
      ep = base + re;
      if (sp >= ep) {
        finish(0, 0, 0);
      } else if (--rem == 0) {
        finish(ep - sp, sp, ep);
      } else if (shortcut && ep - sp <= max_rows && rem >= iv.verify_min && rem <= kVerifyMax && (sp & dense_mask) == 0) {
        // Few rows left: the suffix of row r starts at SA[r], and r stays in the interval iff the
        // rem characters before that text position equal the rest of the pattern. Ask for the
        // suffix-array entries now (lane j: row sp + j), use them next trip.
        CSFM_CHK(ep <= iv.n, "interval inside the suffix array");
        if ((uint32_t)j < ep - sp) vp = iv.dense[(sp >> iv.dense_shift) + j];
        vstage = 1;
        if (kInstr) ++my_checks;
      } else {
        --ptr;
        begin_step(next_byte);
      }
    }
  }
  if (kInstr && a.steps_total) {
    unsigned s = (j == 0) ? my_steps : 0, t = (j == 0) ? my_lookups : 0, u = (j == 0) ? my_checks : 0,
             v = (j == 0) ? my_halves : 0, w = (j == 0) ? my_lines : 0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
      u += __shfl_xor_sync(0xFFFFFFFFu, u, o);
      v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
    }
    if (lane == 0 && w) atomicAdd(a.steps_total + 4, (unsigned long long)w);
    if (lane == 0 && s) atomicAdd(a.steps_total, (unsigned long long)s);
    if (lane == 0 && t) atomicAdd(a.steps_total + 1, (unsigned long long)t);
    if (lane == 0 && u) atomicAdd(a.steps_total + 2, (unsigned long long)u);
    if (lane == 0 && v) atomicAdd(a.steps_total + 3, (unsigned long long)v);
  }
}


// ------------------------------------------------------------------------------------------
// count, first pass of the two-pass form: ONE QUERY PER THREAD.
//
// On an index that carries the text sections almost every query of a realistic batch is "k-mer table lookup,
// one half step (level-1 line), at most four suffix-array entries, text comparison": four dependent fetches
// and nothing else. The sub-warp kernel above spends four lanes and ~90 warp instructions per query on that
// chain and keeps 8 queries per warp in flight (ncu: 7.9 sectors per request, issue slots 57 % busy, 61 % of
// the stall samples waiting on the chain). Here every lane carries its own query through the same stages in
// lock-step with its 31 neighbours: every load instruction fetches 32 different lines (32 sectors per request),
// a warp has 32 chains in flight, and the whole query costs ~10 warp instructions. A query that does not fit
// the pattern (short, long remainder, wide interval, an interval that straddles two lines, a window that
// would wrap around the text) is appended to a.qlist and finished by the sub-warp kernel in a second launch.
// Same arithmetic as count2_kernel<true,.>: fm_index.cpp:79-101 with the text-verification identity of
// csfm_common.cuh.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_u32(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }

// t[0..rem) == p[0..rem), four characters at a time from aligned words (rem <= kVerifyMax). The words that hold
// the last characters may extend up to three bytes past them: the caller guarantees those bytes are readable.
__device__ __forceinline__ bool bytes_equal_global(const uint8_t* __restrict__ t, const uint8_t* __restrict__ p, uint32_t rem) {
  const uint32_t* tw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(t) & ~(uintptr_t)3);
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
  const uint32_t ts = (uint32_t)(reinterpret_cast<uintptr_t>(t) & 3) * 8u, ps = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8u;
  // words needed: those covering [addr, addr + rem)
  const uint32_t tn = (uint32_t)((reinterpret_cast<uintptr_t>(t) & 3) + rem + 3) >> 2;
  const uint32_t pn = (uint32_t)((reinterpret_cast<uintptr_t>(p) & 3) + rem + 3) >> 2;
  uint32_t acc = 0, tlo = __ldg(tw), plo = __ldg(pw);
#pragma unroll 2
  for (uint32_t w = 0; 4u * w < rem; ++w) {
    const uint32_t thi = (w + 1 < tn) ? __ldg(tw + w + 1) : 0u;
    const uint32_t phi = (w + 1 < pn) ? __ldg(pw + w + 1) : 0u;
    const uint32_t x = __funnelshift_r(tlo, thi, ts) ^ __funnelshift_r(plo, phi, ps);
    const uint32_t left = rem - 4u * w;
    const uint32_t valid = left < 4u ? left : 4u;
    acc |= x & __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * valid);
    tlo = thi;
    plo = phi;
  }
  return acc == 0;
}

#ifndef CSFM_COUNT2Q_CTAS
#define CSFM_COUNT2Q_CTAS 4
#endif
// rows verified by the per-thread pass (more go to the second pass) and the lane's private staging slot
constexpr uint32_t kQRows = 2;
struct alignas(16) QSlot {
  uint8_t t[kQRows][48];
  uint8_t p[48];
  uint8_t pad[16];
};
template <bool kInstr>
__global__ void __launch_bounds__(kThreads, CSFM_COUNT2Q_CTAS)
count2q_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ CountArgs a) {
  __shared__ uint4 step_tab[256];
  __shared__ QSlot qslots[kThreads];
  load_step_table(step_tab, iv.hdr);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t kk = iv.kmer_k;
  const bool half = iv.kmer_hi != nullptr;
  const uint32_t npat = (uint32_t)a.npat;
  const uint8_t* const batch_end = a.bytes + a.offs[a.npat];
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t my_lookups = 0, my_halves = 0, my_checks = 0, my_lines = 0;

  // warp-uniform trip count: every lane of a warp runs the same number of rounds (the votes below need all 32)
  for (uint32_t q0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; q0 < npat; q0 += stride) {
    const uint32_t q = q0 + lane;
    const bool live = q < npat;
    bool leave = false;  // this query goes to the second pass
    bool done = false;
    uint32_t cnt = 0;
    uint64_t o0 = 0, o1 = 0;
    if (live) {
      o0 = a.offs[q];
      o1 = a.offs[(uint64_t)q + 1];
    }
    const uint64_t m = o1 - o0;
    const uint32_t used = kk + (half ? 1u : 0u);  // characters consumed before the verification
    if (live && (m < used || m > used + kVerifyMax || kk == 0)) leave = true;
    const bool go = live && !leave;

    // ---- stage 1: key of the last k characters (+ the character in front of them for the half step)
    uint32_t e = 0, mul = 1;
    bool present = true;
    uint4 st = make_uint4(0, 0, 0, 0);
    if (go) {
      for (uint32_t i = 0; i < kk; ++i) {
        const uint4 t = step_tab[a.bytes[o1 - 1 - i]];
        present = present && !(t.y & 0x80000000u);
        e += (t.y & 0xFFu) * mul;
        mul *= iv.kmer_radix;
      }
      if (half) {
        st = step_tab[a.bytes[o1 - 1 - kk]];
        present = present && !(st.y & 0x80000000u);
      }
      if (kInstr) ++my_lookups;
      if (!present) done = true;  // a character that does not occur: count 0 (fm_index.cpp:96)
    }
    // ---- stage 2: table lookup (one 8-byte entry per lane: 32 lines per load instruction)
    uint32_t sp = 0, ep = 0;
    if (go && !done) {
      if (half) {
        const uint2 se = iv.kmer_hi[(size_t)e * 16u + ((st.y & 0xFFu) >> 4)];
        sp = se.x;
        ep = se.y;
      } else if (iv.kmer_tiled) {
        const uint32_t* const t32 = reinterpret_cast<const uint32_t*>(iv.kmer);
        sp = t32[e];
        ep = t32[(size_t)e + 1];
      } else {
        const uint2 se = iv.kmer[e];
        sp = se.x;
        ep = se.y;
      }
      if (sp >= ep) done = true;
    }
    // ---- stage 3 (half-step table): the level-1 half of the pending step, both ends from ONE line.
    // A lane loads only the 32-byte chunks it needs: those with symbols before `oe`, and the one that holds
    // the counter of v (every load instruction costs the L1 one tag cycle per distinct line it touches).
    if (half && go && !done) {
      if (((sp ^ ep) >> 7) != 0u) {
        leave = true;  // the interval straddles two lines
      } else {
        if (kInstr) { ++my_halves; ++my_lines; }
        const uint32_t v = st.y & 15u, vc = v >> 2;
        const uint8_t* const line = iv.levels_last + (sp & ~(kSymsPerLine - 1));  // 128 symbols per 128-byte line
        const uint32_t os = sp & (kSymsPerLine - 1), oe = ep & (kSymsPerLine - 1);
        const uint32_t last = oe >> 5;
        Chunk32 k0 = ldg_nc_v8(line), k1 = chunk_undefined(), k2 = chunk_undefined(), k3 = chunk_undefined();
        if (last >= 1u || vc == 1u) k1 = ldg_nc_v8(line + 32);
        if (last >= 2u || vc == 2u) k2 = ldg_nc_v8(line + 64);
        if (last >= 3u || vc == 3u) k3 = ldg_nc_v8(line + 96);
        const uint32_t c0 = pick4(k0.c0, k0.c1, k0.c2, k0.c3, v), c1 = pick4(k1.c0, k1.c1, k1.c2, k1.c3, v);
        const uint32_t c2 = pick4(k2.c0, k2.c1, k2.c2, k2.c3, v), c3 = pick4(k3.c0, k3.c1, k3.c2, k3.c3, v);
        uint32_t rs = vc == 0u ? c0 : vc == 1u ? c1 : vc == 2u ? c2 : c3, re = rs;
        uint32_t hw = chunk_hits(k0, v);
        rs += (uint32_t)__popc(hw & low_mask((int)os));
        re += (uint32_t)__popc(hw & low_mask((int)oe));
        hw = chunk_hits(k1, v);  // chunks that were not loaded lie at or beyond oe: their mask is empty
        rs += (uint32_t)__popc(hw & low_mask((int)os - 32));
        re += (uint32_t)__popc(hw & low_mask((int)oe - 32));
        hw = chunk_hits(k2, v);
        rs += (uint32_t)__popc(hw & low_mask((int)os - 64));
        re += (uint32_t)__popc(hw & low_mask((int)oe - 64));
        hw = chunk_hits(k3, v);
        rs += (uint32_t)__popc(hw & low_mask((int)os - 96));
        re += (uint32_t)__popc(hw & low_mask((int)oe - 96));
        sp = st.x + rs;  // fm_index.cpp:92-93
        ep = st.x + re;
        if (sp >= ep) done = true;
      }
    }
    // ---- stage 4: what is left of the pattern against the text in front of every row's suffix. The windows
    // (one per row, one of the pattern) are staged in the lane's own shared-memory slot with 16-byte cp.async
    // copies, all in flight together, then compared four characters at a time.
    const uint32_t rem = go ? (uint32_t)(m - used) : 0u;
    if (go && !done && !leave) {
      const uint32_t rows = ep - sp;
      if (rem == 0) {
        cnt = rows;
        done = true;
      } else if (rows > kQRows || rem < iv.verify_min) {
        leave = true;
      } else {
        const uint8_t* const pat = a.bytes + o0;
        const uintptr_t pal = reinterpret_cast<uintptr_t>(pat) & ~(uintptr_t)15;
        const uint32_t np16 = (uint32_t)((reinterpret_cast<uintptr_t>(pat + rem) - pal + 15) >> 4);
        const uint32_t vp0 = iv.dense[sp], vp1 = rows > 1u ? iv.dense[sp + 1] : 0xFFFFFFFFu;
        // an aligned pattern window that would leave the batch, or an occurrence that would wrap around the text
        // start (cyclic BWT): that query keeps stepping in the second pass
        if (reinterpret_cast<const uint8_t*>(pal) + 16 * np16 > batch_end || vp0 < rem || vp1 < rem) {
          leave = true;
        } else {
          if (kInstr) ++my_checks;
          QSlot& sl = qslots[threadIdx.x];
          const uint32_t t0 = (vp0 - rem) & ~15u, t1 = (vp1 - rem) & ~15u;
          const uint32_t n0 = (vp0 - t0 + 15) >> 4, n1 = rows > 1u ? (vp1 - t1 + 15) >> 4 : 0u;
#pragma unroll
          for (uint32_t c = 0; c < 3; ++c) {
            if (c < n0) cp_async16(&sl.t[0][16 * c], iv.text + t0 + 16 * c);
            if (c < n1) cp_async16(&sl.t[1][16 * c], iv.text + t1 + 16 * c);
            if (c < np16) cp_async16(&sl.p[16 * c], reinterpret_cast<const uint8_t*>(pal) + 16 * c);
          }
          cp_async_commit();
          cp_async_wait_all();  // the lane's own copies: nobody else reads its slot
          const uint32_t poff = (uint32_t)(reinterpret_cast<uintptr_t>(pat) & 15u);
          if (windows_equal(sl.t[0], sl.p, (vp0 - rem) & 15u, poff, rem)) ++cnt;
          if (rows > 1u && windows_equal(sl.t[1], sl.p, (vp1 - rem) & 15u, poff, rem)) ++cnt;
          done = true;
        }
      }
    }
    if (live && done && !leave && a.counts) a.counts[q] = cnt;
    // ---- queries for the second pass: one atomic per warp
    const unsigned lv_mask = __ballot_sync(0xFFFFFFFFu, live && leave);
    if (lv_mask) {
      unsigned long long slot0 = 0;
      if (lane == 0) slot0 = atomicAdd(a.qlist_len, (unsigned long long)__popc(lv_mask));
      slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
      if (live && leave) a.qlist[slot0 + __popc(lv_mask & ((1u << lane) - 1u))] = q;
    }
  }
  if (kInstr && a.steps_total) {
    unsigned t = my_lookups, u = my_checks, v = my_halves, w = my_lines;
    for (int o = 16; o > 0; o >>= 1) {
      t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
      u += __shfl_xor_sync(0xFFFFFFFFu, u, o);
      v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
    }
    if (lane == 0 && t) atomicAdd(a.steps_total + 1, (unsigned long long)t);
    if (lane == 0 && u) atomicAdd(a.steps_total + 2, (unsigned long long)u);
    if (lane == 0 && v) atomicAdd(a.steps_total + 3, (unsigned long long)v);
    if (lane == 0 && w) atomicAdd(a.steps_total + 4, (unsigned long long)w);
  }
}

// Fills the k-mer jump table by running the same backward search over every k-symbol string:
// entry e decodes to codes d_0 (LAST character, e % radix), d_1, ... and stores its interval.
__global__ void __launch_bounds__(kThreads)
kmer_build_kernel(const __grid_constant__ IndexView iv, uint2* __restrict__ table, unsigned long long entries, uint32_t k,
                  uint32_t radix) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv_last = iv.levels_last + j * 32;
  const unsigned long long group = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const unsigned long long ngroups = ((unsigned long long)gridDim.x * blockDim.x) >> 2;
  const unsigned long long trips = (entries + ngroups - 1) / ngroups;
  for (unsigned long long t = 0; t < trips; ++t) {
    const unsigned long long e = t * ngroups + group;
    const bool valid = e < entries;
    unsigned long long rest = valid ? e : 0;
    uint32_t byte = tb.byte_of_code[rest % radix];
    rest /= radix;
    uint32_t sp = tb.C[byte], ep = tb.C[byte + 1];  // first step: [C[c], C[c+1])
    bool alive = valid && sp < ep;
    for (uint32_t i = 1; i < k; ++i) {
      const uint32_t code = (uint32_t)(rest % radix);
      rest /= radix;
      byte = tb.byte_of_code[code];
      const bool act = alive && (tb.C[byte + 1] != tb.C[byte]);
      uint32_t rs, re;
      if (two) {
        rank_pair(lv0, code >> 4, sp, ep, act, j, rs, re, iv);
        sp = tb.start1[code >> 4] + rs;
        ep = tb.start1[code >> 4] + re;
      }
      rank_pair(lv_last, code & 15u, sp, ep, act, j, rs, re, iv);
      sp = tb.base_by_code[code] + rs;
      ep = tb.base_by_code[code] + re;
      alive = act && sp < ep;
    }
    if (valid && j == 0) table[e] = alive ? make_uint2(sp, ep) : make_uint2(0u, 0u);
  }
}

// Tiled table, step 1: histogram of the text's cyclic k-grams. Key of the k-gram starting at p: its LAST
// character is the least significant digit (the order count2_kernel builds its key in), so keys sort like
// the k-grams themselves and the exclusive prefix sum of the histogram is sp of every key.
__global__ void __launch_bounds__(kThreads)
kmer_histogram_kernel(const uint8_t* __restrict__ text, unsigned long long n, uint32_t k, uint32_t radix,
                      const BlobHeader* __restrict__ hdr, uint32_t* __restrict__ table) {
  __shared__ uint8_t code_of_byte[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) code_of_byte[i] = hdr->code_of_byte[i];
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    unsigned long long e = 0;
    for (uint32_t i = 0; i < k; ++i) {  // first character first: e = e * radix + code
      unsigned long long q = p + i;
      if (q >= n) q -= n;  // rows are rotations: the k-gram wraps around the text end
      e = e * radix + code_of_byte[text[q]];
    }
    atomicAdd(&table[e], 1u);
  }
}

// Fills the half-step table from the finished k-mer table: entry e * 16 + g = the interval of k-mer e
// mapped through level 0 for the high nibble g. The two lines of an interval are loaded once and
// serve all 16 nibbles.
__global__ void __launch_bounds__(kThreads)
kmer_hi_build_kernel(const __grid_constant__ IndexView iv, const uint2* __restrict__ table, uint2* __restrict__ table_hi,
                     unsigned long long entries, bool tiled) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int j = threadIdx.x & 3;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const unsigned long long group = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const unsigned long long ngroups = ((unsigned long long)gridDim.x * blockDim.x) >> 2;
  const unsigned long long trips = (entries + ngroups - 1) / ngroups;
  for (unsigned long long t = 0; t < trips; ++t) {
    const unsigned long long e = t * ngroups + group;
    const bool valid = e < entries;
    uint2 se = make_uint2(0u, 0u);
    if (valid) se = tiled ? make_uint2(reinterpret_cast<const uint32_t*>(table)[e], reinterpret_cast<const uint32_t*>(table)[e + 1]) : table[e];
    const bool alive = valid && se.x < se.y;
    const uint32_t ls = se.x & ~(kSymsPerLine - 1), le = se.y & ~(kSymsPerLine - 1);
    Chunk32 ks = chunk_undefined(), ke = chunk_undefined();
    if (alive) {
      ks = ldg_nc_v8(lv0 + ls);
      ke = ldg_nc_v8(lv0 + le);
    }
#pragma unroll 1
    for (uint32_t g = 0; g < 16; ++g) {
      const uint32_t rs = group4_sum(chunk_partial(chunk_counter(ks, g, j), chunk_hits(ks, g), se.x - ls, j));
      const uint32_t re = group4_sum(chunk_partial(chunk_counter(ke, g, j), chunk_hits(ke, g), se.y - le, j));
      if (valid && j == 0)
        table_hi[e * 16 + g] = (alive && rs < re) ? make_uint2(tb.start1[g] + rs, tb.start1[g] + re) : make_uint2(0u, 0u);
    }
  }
}

// TMA variant (CSFM_PATTERN_STAGING=tma): every warp stages 32-pattern chunks — their offsets and
// their packed bytes — into a double buffer in shared memory with cp.async.bulk (UBLKCP) signalled
// through mbarriers, one chunk ahead of the one being consumed, and sub-warps copy short patterns to
// a private slot so the buffer can be refilled under them. Measured on B200 (profiles/README.md)
// it is ~15 % slower than the direct variant on C3: the kernel is bound by issue slots and HBM
// fetch latency, pattern bytes are 0.1 % of the traffic, and the staging control costs more issue
// slots than the two dependent loads it removes. Kept selectable and covered by the parity tests.
__global__ void __launch_bounds__(kThreads, 5)
count2_tma_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ CountArgs a) {
  __shared__ Tables tb;
  __shared__ WarpStage stages[kThreads / 32];
  __shared__ __align__(16) uint8_t priv_all[kThreads / 4][kPrivBytes];  // one private pattern slot per sub-warp
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  WarpStage& st = stages[threadIdx.x >> 5];
  uint8_t* const priv = priv_all[threadIdx.x >> 2];
  if (lane == 0) {
    mbar_init(&st.bar_offs[0], 1);
    mbar_init(&st.bar_offs[1], 1);
    mbar_init(&st.bar_bytes[0], 1);
    mbar_init(&st.bar_bytes[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  load_tables(tb, iv.hdr);  // ends with __syncthreads()

  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv_last = iv.levels_last + j * 32;
  const uint32_t npat = (uint32_t)a.npat;  // < 2^32 per launch (count_device slices larger batches)
  // TMA bulk copies need 16-byte aligned sources; otherwise every chunk is read straight from global
  const bool can_stage = ((reinterpret_cast<uintptr_t>(a.bytes) | reinterpret_cast<uintptr_t>(a.offs)) & 15) == 0;

  // ---- warp-uniform staging state: chunk being consumed + chunk being prefetched -------------
  uint32_t cur_next = 0, cur_end = 0, cur_base = 0, cur_buf = 1;
  uint32_t pf_state = 0;  // 0 idle, 1 offsets in flight, 2 chunk ready (bytes in flight / landed / direct)
  uint32_t pf_base = 0, pf_end = 0;
  uint32_t par_offs = 0, par_bytes = 0;  // mbarrier phase parity, bit b = buffer b
  uint32_t direct_mask = 0;              // bit b: chunk in buffer b is read straight from global memory
  bool exhausted = false;

  // ---- per-sub-warp query state ---------------------------------------------------------------
  bool active = false;
  uint32_t q = 0;                // query index
  const uint8_t* ptr = nullptr;  // address of the character being processed (shared or global)
  uint32_t rem = 0;              // characters left including the current one
  uint32_t mybuf = 2;            // staging buffer the pattern bytes live in (2 = global memory)
  uint32_t sp = 0, ep = 0, base = 0, add0 = 0, code = 0, next_byte = 0;
  uint32_t my_steps = 0;

  auto finish = [&](uint32_t cnt, uint32_t lo, uint32_t hi) {
    if (j == 0) {
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
  // Sets up the step that prepends byte b to the interval [sp,ep).
  auto begin_step = [&](uint32_t b) {
    ++my_steps;
    if (tb.C[b + 1] == tb.C[b]) {  // symbol absent: occ(c,.) == 0 -> sp == ep (fm_index.cpp:96)
      finish(0, 0, 0);
      return;
    }
    code = tb.code_of_byte[b];
    base = tb.base_by_byte[b];
    add0 = tb.start1[code >> 4];
    if (rem > 1) next_byte = ptr[-1];  // in flight during the rank levels
  };

  for (;;) {
    // ---- A. pattern staging pipeline: TMA bulk copies into the buffer that is not being consumed
    if (pf_state == 1) {
      const uint32_t b = cur_buf ^ 1u;
      // the bytes of the chunk that used to live in buffer b may still be read by a long query
      if (!__any_sync(0xFFFFFFFFu, active && mybuf == b)) {
        mbar_wait(&st.bar_offs[b], (par_offs >> b) & 1u);
        par_offs ^= 1u << b;
        int direct = 1;
        if (lane == 0) {
          const uint64_t o_first = st.offs[b][0], o_last = st.offs[b][pf_end - pf_base];
          const uint8_t* src = a.bytes + o_first;
          const uint8_t* src_al = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
          const uint64_t span = (uint64_t)(src - src_al) + (o_last - o_first);
          const bool tail_overrun = (pf_end == npat) && ((span & 15) != 0);  // rounding up would read past the batch
          if (o_last != o_first && span <= kStageBytes && !tail_overrun) {
            const uint32_t size = (uint32_t)((span + 15) & ~15ull);
            fence_proxy_async_smem();
            mbar_expect_tx(&st.bar_bytes[b], size);
            tma_bulk_g2s(st.bytes[b], src_al, size, &st.bar_bytes[b]);
            st.base_ptr[b] = st.bytes[b] - (src_al - a.bytes);
            direct = 0;
          } else {
            st.base_ptr[b] = a.bytes;
          }
        }
        direct = __shfl_sync(0xFFFFFFFFu, direct, 0);
        direct_mask = (direct_mask & ~(1u << b)) | ((uint32_t)direct << b);
        __syncwarp();
        pf_state = 2;
      }
    } else if (pf_state == 0 && !exhausted) {
      unsigned long long cbase = 0;
      if (lane == 0) cbase = atomicAdd(a.cursor, (unsigned long long)kChunk);
      cbase = __shfl_sync(0xFFFFFFFFu, cbase, 0);
      if (cbase >= a.npat) {
        exhausted = true;
      } else {
        pf_base = (uint32_t)cbase;
        pf_end = (pf_base + kChunk < npat) ? pf_base + kChunk : npat;
        const uint32_t b = cur_buf ^ 1u;
        const uint32_t ents = (pf_end - pf_base + 2u) & ~1u;  // count+1 offsets, rounded up to 16 bytes
        if (can_stage && (uint64_t)pf_base + ents <= (uint64_t)npat + 1) {
          if (lane == 0) {
            fence_proxy_async_smem();
            mbar_expect_tx(&st.bar_offs[b], ents * 8u);
            tma_bulk_g2s(st.offs[b], a.offs + pf_base, ents * 8u, &st.bar_offs[b]);
          }
          pf_state = 1;
        } else {
          if (lane == 0) st.base_ptr[b] = a.bytes;
          direct_mask |= 1u << b;
          __syncwarp();
          pf_state = 2;
        }
      }
    }

    // ---- B. refill (every sub-warp is at a step boundary here) ------------------------------
    const unsigned need_mask = __ballot_sync(0xFFFFFFFFu, !active) & 0x11111111u;
    if (need_mask) {
      if (cur_next >= cur_end && pf_state == 2) {  // switch to the prefetched chunk
        const uint32_t b = cur_buf ^ 1u;
        if (!((direct_mask >> b) & 1u)) {
          mbar_wait(&st.bar_bytes[b], (par_bytes >> b) & 1u);
          par_bytes ^= 1u << b;
        }
        cur_buf = b;
        cur_base = cur_next = pf_base;
        cur_end = pf_end;
        pf_state = 0;
      }
      const uint32_t avail = cur_end - cur_next;
      const uint32_t my_rank = __popc(need_mask & ((1u << (lane & ~3)) - 1u));
      const uint32_t cnt = __popc(need_mask);
      const bool take = !active && my_rank < avail;
      uint64_t o0 = 0, o1 = 0;
      const uint8_t* pb = nullptr;
      if (take) {
        q = cur_next + my_rank;
        const bool dir = (direct_mask >> cur_buf) & 1u;
        if (dir) {
          o0 = a.offs[q];
          o1 = a.offs[q + 1];
        } else {
          o0 = st.offs[cur_buf][q - cur_base];
          o1 = st.offs[cur_buf][q - cur_base + 1];
        }
        pb = st.base_ptr[cur_buf];
        mybuf = dir ? 2u : cur_buf;
        if (!dir && o1 - o0 <= kPrivBytes) {
          // Short pattern of a staged chunk: move it to the sub-warp's private slot so that the
          // chunk buffer can be refilled by TMA while this query is still running.
          for (uint32_t k = j; k < (uint32_t)(o1 - o0); k += 4) priv[k] = pb[o0 + k];
          pb = priv - o0;
          mybuf = 2u;
        }
      }
      __syncwarp();  // the private-slot bytes a lane reads below were written by its three neighbours
      if (take) {
        const uint64_t m = o1 - o0;
        active = true;
        if (m == 0) {
          // count("") == n (fm_index.cpp:80); locate("") is empty (fm_index.cpp:109)
          if (j == 0) {
            if (a.counts) a.counts[q] = iv.n;
            if (a.sp_ep) { a.sp_ep[2 * (uint64_t)q] = 0; a.sp_ep[2 * (uint64_t)q + 1] = 0; }
            if (a.row_sp) { a.row_sp[q] = 0; a.row_cnt[q] = 0; }
          }
          active = false;
        } else {
          // first step needs no rank: occ(c,0) = 0 and occ(c,n) = freq[c]  => [C[c], C[c+1])
          const uint32_t b = pb[o1 - 1];
          sp = tb.C[b];
          ep = tb.C[b + 1];
          ++my_steps;
          if (sp >= ep) {
            finish(0, 0, 0);
          } else if (m == 1) {
            finish(ep - sp, sp, ep);
          } else {
            rem = (uint32_t)(m - 1);
            ptr = pb + (o1 - 2);
            begin_step(*ptr);
          }
        }
      }
      cur_next += (cnt < avail) ? cnt : avail;
    }
    if (exhausted && pf_state == 0 && cur_next >= cur_end && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- C. one backward-search step: sp/ep <- base[c] + rank_last(lo, start1[hi] + rank_0(hi, .))
    uint32_t rs, re;
    if (two) {
      rank_pair(lv0, code >> 4, sp, ep, active, j, rs, re, iv);
      sp = add0 + rs;
      ep = add0 + re;
    }
    rank_pair(lv_last, code & 15u, sp, ep, active, j, rs, re, iv);
    if (active) {
      sp = base + rs;  // fm_index.cpp:92-93
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
  if (a.steps_total) {
    unsigned s = (j == 0) ? my_steps : 0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0 && s) atomicAdd(a.steps_total, (unsigned long long)s);
  }
}

// ------------------------------------------------------------------------------------------
// count for ONE pattern: one warp, the pattern (as per-character step entries) in the kernel parameters, the result
// (count, interval, sequence number) written to mapped pinned host memory in one 16-byte store the host is spinning on.
// Same search as count2_kernel<false,.>: k-mer table for the last k characters, then one rank step per
// character (fm_index.cpp:79-101).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
count_single2_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ SingleQuery sq, SingleResult* __restrict__ res) {
  const int lane = threadIdx.x;
  const int j = lane & 3;
  const uint32_t m = sq.len;
  const bool two = iv.L == 2;
  const uint4* const ent = sq.ent;  // per pattern character, in the kernel parameters: no table fetch before the first rank
  const uint32_t c_after_last = sq.c_after_last;
  uint32_t sp = 0, ep = 0;
  unsigned long long cnt = 0;
  if (m == 0) {
    cnt = iv.n;  // count("") == n (fm_index.cpp:80)
  } else {
    int i;  // index of the next character to prepend
    const uint32_t kk = iv.kmer_k;
    bool alive = true;
    if (kk != 0 && m >= kk) {
      uint32_t e = 0, mul = 1;
      for (uint32_t t = 0; t < kk; ++t) {
        const uint4 x = ent[m - 1 - t];
        alive = alive && !(x.y & 0x80000000u);
        e += (x.y & 0xFFu) * mul;
        mul *= iv.kmer_radix;
      }
      if (alive) {
        if (iv.kmer_tiled) {
          const uint32_t* const t32 = reinterpret_cast<const uint32_t*>(iv.kmer);
          sp = t32[e];
          ep = t32[(size_t)e + 1];
        } else {
          const uint2 se = iv.kmer[e];
          sp = se.x;
          ep = se.y;
        }
      }
      i = (int)m - 1 - (int)kk;
    } else {
      const uint4 x = ent[m - 1];
      sp = x.w;  // first step needs no rank: [C[c], C[c+1])
      ep = c_after_last;
      i = (int)m - 2;
    }
    alive = alive && sp < ep;
    const uint8_t* const lv0 = iv.levels + j * 32;
    const uint8_t* const lv_last = iv.levels_last + j * 32;
    const bool mine = lane < 4;
    for (; i >= 0 && alive; --i) {  // warp-uniform: every lane holds the same sp/ep after the shuffles of its group
      const uint4 x = ent[i];
      if (x.y & 0x80000000u) { alive = false; break; }
      uint32_t rs, re, s1 = sp, e1 = ep;
      if (two) {
        rank_pair(lv0, (x.y & 0xFFu) >> 4, sp, ep, mine, j, rs, re, iv);
        s1 = x.z + rs;
        e1 = x.z + re;
      }
      rank_pair(lv_last, x.y & 15u, s1, e1, mine, j, rs, re, iv);
      sp = __shfl_sync(0xFFFFFFFFu, x.x + rs, 0);
      ep = __shfl_sync(0xFFFFFFFFu, x.x + re, 0);
      alive = sp < ep;
    }
    if (alive) cnt = ep - sp;
    else sp = ep = 0;
  }
  if (lane == 0) {
    // one 16-byte store to the mapped pinned slot: a single write transaction on the link, so the host that sees the
    // new sequence number sees the result with it (no system-scope fence between result and flag)
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(res), "r"((uint32_t)cnt), "r"(sp), "r"(ep), "r"(sq.seq) : "memory");
  }
}

// One level of access(p) fused with rank (both read the same line): returns the symbol at p in
// `v` and rank_l(v, p) as the function value. All 32 lanes must call it (shuffles).
__device__ __forceinline__ uint32_t access_rank_level(const uint8_t* __restrict__ lv, uint32_t p, bool active, int lane,
                                                      int j, uint32_t& v, const IndexView& iv) {
  Chunk32 k = chunk_undefined();
  const uint32_t line = p & ~(kSymsPerLine - 1), off = p - line;
  if (active) { check_line(iv, lv + line, 32); k = ldg_nc_v8(lv + line); }
  v = __shfl_sync(0xFFFFFFFFu, chunk_symbol(k, off), (lane & ~3) | (int)(off >> 5));
  return group4_sum(chunk_partial(chunk_counter(k, v, j), chunk_hits(k, v), off, j));
}

// ------------------------------------------------------------------------------------------
// locate: rows -> text positions (fm_index.cpp:125-153, LF of fm_index.hpp:62-66)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 8)
walk2_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ WalkArgs a) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv1 = iv.levels_last + j * 32;
  WarpQueue wq;

  bool active = false;
  unsigned long long slot = 0;
  uint32_t start = 0, p = 0, steps = 0;
  uint32_t my_lf = 0;

  // Where the reference throws, the whole query fails: attribute the slot to its query.
  auto fail_walk = [&](int why) {
    if (j == 0) {
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
  auto emit = [&](uint32_t row) {  // row is sampled: SA[row] = ssa[row/stride]
    const uint32_t k = sample_index(iv, row);
    if (k >= iv.nsamp) {  // fm_index.cpp:141-146 (unreachable for a consistent index)
      fail_walk((int)CSFM_Q_SSA_OOB);
      return;
    }
    if (j == 0) {
      uint64_t pos = (uint64_t)iv.ssa[k] + steps;  // fm_index.cpp:147-152
      if (pos >= iv.n) pos -= iv.n;                // sa_val < n and steps < n
      a.out_pos[slot] = pos;
    }
    active = false;
  };

  for (;;) {
    const unsigned long long item = queue_take(wq, !active, lane, a.cursor, a.total);
    if (item != ~0ull) {
      slot = a.first + item;
      start = a.rows_implicit ? a.row_base + (uint32_t)slot : (uint32_t)a.out_pos[slot];
      steps = 0;
      active = true;
      p = start;
      if (row_is_sampled(iv, start)) emit(start);
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- one LF step: LF(i) = C[c] + occ(c,i) = base[c] + rank_last(lo, .)  (fm_index.hpp:62-66)
    uint32_t v, code = 0;
    uint32_t r = access_rank_level(lv0, p, active, lane, j, v, iv);
    if (two) {
      code = v << 4;
      r = access_rank_level(lv1, tb.start1[v] + r, active, lane, j, v, iv);
    }
    code |= v;
    if (active) {
      const uint32_t row = tb.base_by_code[code] + r;
      ++steps;
      ++my_lf;
      if (row_is_sampled(iv, row)) {
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
    unsigned s = (j == 0) ? my_lf : 0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0 && s) atomicAdd(a.lf_total, (unsigned long long)s);
  }
}

// ------------------------------------------------------------------------------------------
// untext: the text back out of the index (cs::FMIndex::extract on an index that was loaded without its TEXT
// section, fm_index.cpp:163-167). One sub-warp per SAMPLED row k: it starts at text position ssa[k] and walks LF
// until the next sampled row, writing BWT[row] = T[position - 1] at every step. LF is a permutation, so the walks
// of all sampled rows together visit every row exactly once (rows on an LF cycle without a sampled row — texts
// whose locate() throws — are never reached: the caller compares the number of bytes written with n).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 8)
untext2_kernel(const __grid_constant__ IndexView iv, uint8_t* __restrict__ out, unsigned long long* __restrict__ cursor,
               unsigned long long* __restrict__ written) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv1 = iv.levels_last + j * 32;
  WarpQueue wq;
  bool active = false;
  uint32_t p = 0, pos = 0, steps = 0, my_written = 0;
  for (;;) {
    const unsigned long long item = queue_take(wq, !active, lane, cursor, iv.nsamp);
    if (item != ~0ull) {
      p = (uint32_t)item * iv.stride;
      pos = iv.ssa[item];
      steps = 0;
      active = true;
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;
    uint32_t v, code = 0;
    uint32_t r = access_rank_level(lv0, p, active, lane, j, v, iv);
    if (two) {
      code = v << 4;
      r = access_rank_level(lv1, tb.start1[v] + r, active, lane, j, v, iv);
    }
    code |= v;
    if (active) {
      pos = pos == 0 ? iv.n - 1 : pos - 1;  // BWT[row] = T[(SA[row] - 1) mod n]  (bwt.hpp:10-13)
      if (j == 0) out[pos] = tb.byte_of_code[code];
      ++my_written;
      p = tb.base_by_code[code] + r;
      if (row_is_sampled(iv, p) || ++steps >= iv.n) active = false;
    }
  }
  unsigned s = (j == 0) ? my_written : 0;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0 && s) atomicAdd(written, (unsigned long long)s);
}

// ------------------------------------------------------------------------------------------
// access: BWT[i] for all i (wavelet.cpp:102-128) — verification / export, not a query path
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
access2_kernel(const __grid_constant__ IndexView iv, uint8_t* __restrict__ out) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const bool two = iv.L == 2;
  const uint8_t* const lv0 = iv.levels + j * 32;
  const uint8_t* const lv1 = iv.levels_last + j * 32;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 2;
  const uint64_t trips = ((uint64_t)iv.n + ngroups - 1) / ngroups;
  for (uint64_t t = 0; t < trips; ++t) {
    const uint64_t i = t * ngroups + group;
    const bool valid = i < iv.n;
    uint32_t v;
    uint32_t r = access_rank_level(lv0, valid ? (uint32_t)i : 0u, valid, lane, j, v, iv);
    uint32_t code = v;
    if (two) {
      const uint32_t hi = v;
      r = access_rank_level(lv1, tb.start1[hi] + r, valid, lane, j, v, iv);
      code = (hi << 4) | v;
    }
    if (valid && j == 0) out[i] = tb.byte_of_code[code];
  }
}

int blocks_per_sm(const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm;
}

}  // namespace

// The verification variant needs the text sections, no interval outputs and 16-byte aligned bytes.
static bool use_shortcut(const IndexView& iv, const CountArgs& a) {
  return iv.text != nullptr && a.sp_ep == nullptr && a.row_sp == nullptr &&
         (reinterpret_cast<uintptr_t>(a.bytes) & 15) == 0;
}

// First pass of the two-pass count: needs the text sections with the full suffix array, a k-mer table, no interval
// outputs. (Unlike the sub-warp verification it does not need the batch bytes to be 16-byte aligned.)
bool count2q_eligible(const IndexView& iv, const CountArgs& a) {
  return iv.text != nullptr && iv.dense != nullptr && iv.dense_shift == 0 && iv.kmer_k != 0 && iv.kmer != nullptr &&
         a.sp_ep == nullptr && a.row_sp == nullptr && a.counts != nullptr;
}
void launch_count2q(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream) {
  if (a.steps_total) count2q_kernel<true><<<grid, kThreads, 0, stream>>>(iv, a);
  else count2q_kernel<false><<<grid, kThreads, 0, stream>>>(iv, a);
}
int max_blocks_per_sm_count2q(const CountArgs& a) {
  return a.steps_total ? blocks_per_sm((const void*)count2q_kernel<true>) : blocks_per_sm((const void*)count2q_kernel<false>);
}

void launch_count2(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream, bool tma_staging) {
  if (tma_staging)
    count2_tma_kernel<<<grid, kThreads, 0, stream>>>(iv, a);
  else if (use_shortcut(iv, a)) {
    if (a.steps_total) count2_kernel<true, true><<<grid, kThreads, 0, stream>>>(iv, a);
    else count2_kernel<true, false><<<grid, kThreads, 0, stream>>>(iv, a);
  } else {
    if (a.steps_total) count2_kernel<false, true><<<grid, kThreads, 0, stream>>>(iv, a);
    else count2_kernel<false, false><<<grid, kThreads, 0, stream>>>(iv, a);
  }
}
int build_kmer_table(csfm_index* idx, cudaStream_t stream) {
  const BlobHeader& h = idx->h;
  if (!h.kmer_k) return CSFM_OK;
  unsigned long long entries = 1;
  for (uint32_t i = 0; i < h.kmer_k; ++i) entries *= h.kmer_radix;
  uint2* table = reinterpret_cast<uint2*>(idx->d_blob + h.off_kmer);
  const unsigned long long want = (entries * 4 + kThreads - 1) / kThreads;
  const int grid = (int)std::min<unsigned long long>(want, (unsigned long long)idx->num_sms * blocks_per_sm((const void*)kmer_build_kernel));
  IndexView v = idx->view;  // the table is being written: the builder itself must not consult it
  v.kmer = nullptr;
  v.kmer_hi = nullptr;
  v.kmer_k = 0;
  if (h.kmer_tiled) {
    // histogram of the text's k-grams (the text copy inside the blob), then an in-place exclusive prefix sum
    // over entries + 1 counters: table[e] = sp(e), table[entries] = n
    if (!h.off_text) return fail(CSFM_ERR_FORMAT, "tiled k-mer table without text section");
    uint32_t* t32 = reinterpret_cast<uint32_t*>(table);
    CSFM_CUDA(cudaMemsetAsync(t32, 0, (entries + 1) * 4, stream));
    kmer_histogram_kernel<<<idx->num_sms * 8, kThreads, 0, stream>>>(idx->d_blob + h.off_text, h.n, h.kmer_k, h.kmer_radix,
                                                                  v.hdr, t32);
    CSFM_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    CSFM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, t32, t32, (int64_t)(entries + 1), stream));
    void* d_tmp = nullptr;
    CSFM_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
    const cudaError_t se = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, t32, t32, (int64_t)(entries + 1), stream);
    const cudaError_t sy = cudaStreamSynchronize(stream);
    cudaFree(d_tmp);
    CSFM_CUDA(se);
    CSFM_CUDA(sy);
  } else {
    kmer_build_kernel<<<grid, kThreads, 0, stream>>>(v, table, entries, h.kmer_k, h.kmer_radix);
    CSFM_CUDA(cudaGetLastError());
  }
  if (h.off_kmer_hi) {
    uint2* table_hi = reinterpret_cast<uint2*>(idx->d_blob + h.off_kmer_hi);
    const int grid_hi = (int)std::min<unsigned long long>(want, (unsigned long long)idx->num_sms * blocks_per_sm((const void*)kmer_hi_build_kernel));
    kmer_hi_build_kernel<<<grid_hi, kThreads, 0, stream>>>(v, table, table_hi, entries, h.kmer_tiled != 0);
    CSFM_CUDA(cudaGetLastError());
  }
  CSFM_CUDA(cudaStreamSynchronize(stream));
  return CSFM_OK;
}

void launch_count_single2(const IndexView& iv, const SingleQuery& q, SingleResult* d_result, cudaStream_t stream) {
  count_single2_kernel<<<1, 32, 0, stream>>>(iv, q, d_result);
}
void launch_walk2(const IndexView& iv, const WalkArgs& a, int grid, cudaStream_t stream) {
  walk2_kernel<<<grid, kThreads, 0, stream>>>(iv, a);
}
void launch_untext2(const IndexView& iv, uint8_t* out, unsigned long long* cursor, unsigned long long* written, int num_sms,
                    cudaStream_t stream) {
  const unsigned long long want = ((unsigned long long)iv.nsamp * 4 + kThreads - 1) / kThreads;
  const int grid = (int)std::min<unsigned long long>(std::max<unsigned long long>(want, 1), (unsigned long long)num_sms * blocks_per_sm((const void*)untext2_kernel));
  untext2_kernel<<<grid, kThreads, 0, stream>>>(iv, out, cursor, written);
}
void launch_access2(const IndexView& iv, uint8_t* out, int grid, cudaStream_t stream) {
  access2_kernel<<<grid, kThreads, 0, stream>>>(iv, out);
}
int max_blocks_per_sm_count2(bool tma_staging, const IndexView& iv, const CountArgs& a) {
  if (tma_staging) return blocks_per_sm((const void*)count2_tma_kernel);
  return use_shortcut(iv, a) ? blocks_per_sm((const void*)count2_kernel<true, false>)
                             : blocks_per_sm((const void*)count2_kernel<false, false>);
}
int max_blocks_per_sm_walk2() { return blocks_per_sm((const void*)walk2_kernel); }
int max_blocks_per_sm_access2() { return blocks_per_sm((const void*)access2_kernel); }

}  // namespace csfm
