// csfm_query.cu — backward-search (count), LF-walk (locate) and access kernels for sm_100a.
//
// Replaces cs::FMIndex::count / locate (/root/reference/src/api/fm_index.cpp:79-157),
// cs::WaveletTree::rank / access (src/core/wavelet.cpp:59-128) and cs::BitVector::rank1
// (src/core/bitvector.cpp:165-230) on the query path.
//
// Execution model (all three kernels): a sub-warp of 4 lanes owns one query / occurrence /
// position. One rank = one 64-byte line = four coalesced 128-bit loads (one per lane), masked
// popc, two xor-shuffles. The 8 sub-warps of a warp run ONE flat state machine in lock-step —
// each loop trip is "one wavelet level for whatever query my sub-warp currently holds" — so
// queries of different length never diverge the instruction stream, and a sub-warp that
// finishes refills from a warp-local chunk of the batch cursor (ballot-ranked, one global
// atomic per 32 queries). Grids are persistent: SM count x resident CTAs.
#include <algorithm>
#include <cstring>

#include <cub/device/device_scan.cuh>

#include "csfm_dna.cuh"
#include "csfm_host.hpp"
#include "csfm_kernels.cuh"

namespace csfm {

namespace {

// ------------------------------------------------------------------------------------------
// count: backward search (fm_index.cpp:79-101), layout 1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 6)
count_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ CountArgs a) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);

  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const uint32_t L = iv.L;
  WarpQueue wq;

  bool active = false;
  unsigned long long q = 0;       // query index
  const uint8_t* ptr = nullptr;   // address of the character being processed
  uint32_t rem = 0;               // characters left including the current one
  uint32_t sp_pos = 0, ep_pos = 0, base = 0, code = 0, level = 0, next_byte = 0;
  uint32_t my_steps = 0;

  auto finish = [&](uint32_t cnt, uint32_t sp, uint32_t ep) {
    if (j == 0) {
      if (a.counts) a.counts[q] = cnt;
      if (a.sp_ep) {
        a.sp_ep[2 * q] = sp;
        a.sp_ep[2 * q + 1] = ep;
      }
      if (a.row_sp) {
        a.row_sp[q] = sp;
        a.row_cnt[q] = cnt < a.limit32 ? cnt : a.limit32;
      }
    }
    active = false;
  };
  // Begin the step for byte `b` on interval [sp,ep): sets the per-step constants.
  auto begin_step = [&](uint32_t b, uint32_t sp, uint32_t ep) {
    ++my_steps;
    if (tb.C[b + 1] == tb.C[b]) {  // symbol absent: occ(c,.) == 0 -> sp == ep (fm_index.cpp:96)
      finish(0, 0, 0);
      return;
    }
    code = tb.code_of_byte[b];
    base = tb.base_by_byte[b];
    sp_pos = sp;
    ep_pos = ep;
    level = 0;
    if (rem > 1) next_byte = ptr[-1];  // prefetch: in flight during the L rank levels
  };

  for (;;) {
    // ---- refill -----------------------------------------------------------------------
    const unsigned long long item = queue_take(wq, !active, lane, a.cursor, a.npat);
    if (item != ~0ull) {
      q = item;
      const uint64_t o0 = a.offs[q], o1 = a.offs[q + 1];
      const uint64_t m = o1 - o0;
      active = true;
      if (m == 0) {
        // count("") == n (fm_index.cpp:80); locate("") is empty (fm_index.cpp:109)
        if (j == 0) {
          if (a.counts) a.counts[q] = iv.n;
          if (a.sp_ep) { a.sp_ep[2 * q] = 0; a.sp_ep[2 * q + 1] = 0; }
          if (a.row_sp) { a.row_sp[q] = 0; a.row_cnt[q] = 0; }
        }
        active = false;
      } else {
        // first step needs no rank: occ(c,0) = 0 and occ(c,n) = freq[c]  => [C[c], C[c+1])
        const uint32_t b = a.bytes[o1 - 1];
        const uint32_t sp = tb.C[b], ep = tb.C[b + 1];
        ++my_steps;
        if (sp >= ep) {
          finish(0, 0, 0);
        } else if (m == 1) {
          finish(ep - sp, sp, ep);
        } else {
          rem = (uint32_t)(m - 1);
          ptr = a.bytes + (o1 - 2);
          begin_step(*ptr, sp, ep);
        }
      }
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // ---- one wavelet level for the two interval ends ------------------------------------
    uint4 ws = make_uint4(0, 0, 0, 0), we = ws;
    uint32_t os = 0, oe = 0;
    if (active) {
      const uint8_t* lv = iv.levels + (uint64_t)level * iv.level_stride + j * 16;
      const uint32_t bs = sp_pos / kPayloadBits, be = ep_pos / kPayloadBits;
      os = sp_pos - bs * kPayloadBits;
      oe = ep_pos - be * kPayloadBits;
      ws = ldg_nc_v4(lv + (uint64_t)bs * kLineBytes);
      we = (be == bs) ? ws : ldg_nc_v4(lv + (uint64_t)be * kLineBytes);
    }
    const uint32_t rs = group4_sum(lane_partial_rank(ws, os, j));
    const uint32_t re = group4_sum(lane_partial_rank(we, oe, j));
    if (active) {
      const uint32_t bit = (code >> (L - 1 - level)) & 1u;
      const uint32_t z = iv.zeros[level];
      sp_pos = bit ? z + rs : sp_pos - rs;  // wavelet.cpp:73-87 for the `end` position
      ep_pos = bit ? z + re : ep_pos - re;
      ++level;
      if (level == L) {
        const uint32_t sp = base + sp_pos, ep = base + ep_pos;  // fm_index.cpp:92-93
        if (sp >= ep) {
          finish(0, 0, 0);
        } else if (--rem == 0) {
          finish(ep - sp, sp, ep);
        } else {
          --ptr;
          begin_step(next_byte, sp, ep);
        }
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
// locate: rows -> text positions (fm_index.cpp:125-153, LF of fm_index.hpp:62-66)
// ------------------------------------------------------------------------------------------
// out_pos[out_offs[q] + k] = sp[q] + k  (SA rows of query q, in row order)
__global__ void expand_rows_kernel(const uint32_t* __restrict__ row_sp,
                                   const uint64_t* __restrict__ out_offs, uint64_t npat,
                                   uint64_t* __restrict__ out_pos) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t q = warp; q < npat; q += nwarps) {
    const uint64_t o0 = out_offs[q], o1 = out_offs[q + 1];
    const uint64_t sp = row_sp[q];
    for (uint64_t k = lane; k < o1 - o0; k += 32) out_pos[o0 + k] = sp + k;
  }
}

__global__ void __launch_bounds__(kThreads, 6)
walk_kernel(const __grid_constant__ IndexView iv, const __grid_constant__ WalkArgs a) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const uint32_t L = iv.L;
  WarpQueue wq;

  bool active = false;
  unsigned long long slot = 0;
  uint32_t start = 0, p = 0, steps = 0, code = 0, level = 0;
  uint32_t my_lf = 0;

  // Where the reference throws, the whole query fails: attribute the slot to its query.
  auto fail_walk = [&](int why) {
    if (j == 0) {
      unsigned long long lo = 0, hi = a.npat;  // last q with out_offs[q] <= slot
      while (hi - lo > 1) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (a.out_offs[mid] <= slot) lo = mid; else hi = mid;
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
      if (row_is_sampled(iv, start)) {
        emit(start);
      } else {
        p = start;
        code = 0;
        level = 0;
      }
    }
    if (wq.exhausted && !__any_sync(0xFFFFFFFFu, active)) break;

    // one level of access(p) fused with rank: both follow the same line (wavelet.cpp:102-128)
    uint4 w = make_uint4(0, 0, 0, 0);
    uint32_t off = 0;
    if (active) {
      const uint32_t blk = p / kPayloadBits;
      off = p - blk * kPayloadBits;
      w = ldg_nc_v4(iv.levels + (uint64_t)level * iv.level_stride + (uint64_t)blk * kLineBytes + j * 16);
    }
    const uint32_t r = group4_sum(lane_partial_rank(w, off, j));
    const uint32_t kw = 1 + (off >> 5);  // line word holding bit `off`
    const uint32_t comp = kw & 3;
    const uint32_t mine = comp == 0 ? w.x : comp == 1 ? w.y : comp == 2 ? w.z : w.w;
    const uint32_t wsel = __shfl_sync(0xFFFFFFFFu, mine, (lane & ~3) | (kw >> 2));
    if (active) {
      const uint32_t bit = (wsel >> (off & 31)) & 1u;
      p = bit ? iv.zeros[level] + r : p - r;
      code = (code << 1) | bit;
      ++level;
      if (level == L) {
        // p = path(c,i); LF(i) = C[c] + rank(c,i) = base[c] + path(c,i)  (fm_index.hpp:62-66)
        const uint32_t row = tb.base_by_code[code] + p;
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
          code = 0;
          level = 0;
        }
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
// access: BWT[i] for all i (wavelet.cpp:102-128) — verification / export, not a query path
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
access_kernel(const __grid_constant__ IndexView iv, uint8_t* __restrict__ out) {
  __shared__ Tables tb;
  load_tables(tb, iv.hdr);
  const int lane = threadIdx.x & 31;
  const int j = lane & 3;
  const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 2;
  const uint64_t trips = ((uint64_t)iv.n + ngroups - 1) / ngroups;
  for (uint64_t t = 0; t < trips; ++t) {
    const uint64_t i = t * ngroups + group;
    const bool valid = i < iv.n;
    uint32_t p = valid ? (uint32_t)i : 0, code = 0;
    for (uint32_t level = 0; level < iv.L; ++level) {
      const uint32_t blk = p / kPayloadBits, off = p - blk * kPayloadBits;
      const uint4 w = ldg_nc_v4(iv.levels + (uint64_t)level * iv.level_stride +
                                (uint64_t)blk * kLineBytes + j * 16);
      const uint32_t r = group4_sum(lane_partial_rank(w, off, j));
      const uint32_t kw = 1 + (off >> 5), comp = kw & 3;
      const uint32_t mine = comp == 0 ? w.x : comp == 1 ? w.y : comp == 2 ? w.z : w.w;
      const uint32_t wsel = __shfl_sync(0xFFFFFFFFu, mine, (lane & ~3) | (kw >> 2));
      const uint32_t bit = (wsel >> (off & 31)) & 1u;
      p = bit ? iv.zeros[level] + r : p - r;
      code = (code << 1) | bit;
    }
    if (valid && j == 0) out[i] = tb.byte_of_code[code];
  }
}

int persistent_grid(const csfm_index* idx, const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return idx->num_sms * per_sm;
}

}  // namespace

unsigned long long* next_counter_slot(csfm_index* idx) {
  const uint32_t s = idx->counter_slot++ % kCounterSlots;
  return idx->d_counters + (size_t)s * kCounterWords;
}

int count_device(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat,
                 uint64_t* d_counts, uint64_t* d_sp_ep, uint32_t* d_row_sp, uint32_t* d_row_cnt,
                 uint64_t limit, cudaStream_t stream) {
  if (npat == 0) return CSFM_OK;
  // the kernels index queries with 32 bits: slice larger batches (offsets stay absolute)
  constexpr uint64_t kMaxPerLaunch = 1ull << 31;
  if (npat > kMaxPerLaunch) {
    if (idx->instr_mask & 1u) return fail(CSFM_ERR_INVALID, "step instrumentation needs batches of at most 2^31 patterns");
    for (uint64_t done = 0; done < npat; done += kMaxPerLaunch) {
      const uint64_t now = std::min(kMaxPerLaunch, npat - done);
      int rc = count_device(idx, d_bytes, d_offs + done, now, d_counts ? d_counts + done : nullptr,
                            d_sp_ep ? d_sp_ep + 2 * done : nullptr, d_row_sp ? d_row_sp + done : nullptr,
                            d_row_cnt ? d_row_cnt + done : nullptr, limit, stream);
      if (rc) return rc;
    }
    return CSFM_OK;
  }
  unsigned long long* ctr = next_counter_slot(idx);
  CSFM_CUDA(cudaMemsetAsync(ctr, 0, kCounterWords * sizeof(unsigned long long), stream));
  CountArgs a{};
  a.bytes = d_bytes;
  a.offs = d_offs;
  a.npat = npat;
  a.counts = d_counts;
  a.sp_ep = d_sp_ep;
  a.row_sp = d_row_sp;
  a.row_cnt = d_row_cnt;
  a.limit32 = limit > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)limit;
  a.cursor = ctr;
  a.steps_total = (idx->instr_mask & 1u) ? ctr + 1 : nullptr;
  const bool nib = idx->view.layout == kLayoutNibble128;
  const bool dna = idx->view.layout == kLayoutDna64;
  const bool tma = idx->tma_staging;
  // layout 3: one lane per query while what the search touches (level lines + k-mer table) lives in the L2 — measured on
  // C2 (56 MB): 1.07e10 against 0.81e10 q/s — a two-lane sub-warp (one 64-byte request per line) when the lines come out
  // of HBM — C5: 5.0e9 against 3.3e9 q/s (tools/ab_count3_lanes.sh, tools/ab_lanes_ctas.sh)
  const uint64_t count_set = (uint64_t)idx->view.L * idx->view.level_stride + idx->view.kmer_entries * sizeof(uint2);
  const int lanes3 = idx->count3_lanes ? idx->count3_lanes : (count_set <= (96ull << 20) ? 1 : 2);
  const int grid_max = dna ? idx->num_sms * max_blocks_per_sm_count3(idx->view, a, lanes3)
                           : nib ? idx->num_sms * max_blocks_per_sm_count2(tma, idx->view, a) : persistent_grid(idx, (const void*)count_kernel);
  const uint64_t want = (npat * (dna ? lanes3 : 4) + kThreads - 1) / kThreads;
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)grid_max);
  const bool timed = (idx->instr_mask & 2u) != 0;
  // Two-pass form on an index with text sections: one query per thread for everything that finishes in "lookup,
  // half step, verification", then the sub-warp kernel over the queries that pass left in the overflow list.
  bool two_pass = nib && !tma && !idx->no_two_pass && count2q_eligible(idx->view, a);
  if (two_pass && idx->two_pass_skip > 0) {
    --idx->two_pass_skip;
    two_pass = false;
  }
  if (two_pass) {
    csfm_index::QListSlot& qs = idx->qslot[idx->qslot_next++ % 8];
    unsigned long long* h_over = static_cast<unsigned long long*>(idx->h_pinned) + 32 + (&qs - idx->qslot);  // bytes 256..319
    if (!qs.done) CSFM_CUDA(cudaEventCreateWithFlags(&qs.done, cudaEventDisableTiming));
    if (qs.npat) {
      if (cudaEventQuery(qs.done) == cudaSuccess) {
        // how the launch that used this slot last went: mostly overflow => this workload is not the common case
        // the first pass is made for; skip it for a while and probe again later
        if (*h_over * 4 > qs.npat) idx->two_pass_skip = 64;
      } else {
        (void)cudaGetLastError();
      }
      CSFM_CUDA(cudaStreamWaitEvent(stream, qs.done, 0));
    }
    int rc = qs.buf.ensure(npat * 4 + 256);
    if (rc) return rc;
    qs.npat = npat;
    a.qlist = qs.buf.as<uint32_t>();
    a.qlist_len = ctr + 6;
    if (timed) CSFM_CUDA(cudaEventRecord(idx->ev0, stream));
    const int grid1 = (int)std::min<uint64_t>((npat + kThreads - 1) / kThreads, (uint64_t)idx->num_sms * max_blocks_per_sm_count2q(a));
    launch_count2q(idx->view, a, grid1, stream);
    CountArgs b = a;
    b.cursor = ctr + 7;
    launch_count2(idx->view, b, idx->num_sms * max_blocks_per_sm_count2(false, idx->view, b), stream, false);
    if (timed) CSFM_CUDA(cudaEventRecord(idx->ev1, stream));
    CSFM_CUDA(cudaGetLastError());
    CSFM_CUDA(cudaMemcpyAsync(h_over, ctr + 6, 8, cudaMemcpyDeviceToHost, stream));
    CSFM_CUDA(cudaEventRecord(qs.done, stream));
    idx->stats.kernel_launches += 2;
  } else {
    if (timed) CSFM_CUDA(cudaEventRecord(idx->ev0, stream));
    if (dna)
      launch_count3(idx->view, a, grid, stream, lanes3);
    else if (nib)
      launch_count2(idx->view, a, grid, stream, tma);
    else
      count_kernel<<<grid, kThreads, 0, stream>>>(idx->view, a);
    if (timed) CSFM_CUDA(cudaEventRecord(idx->ev1, stream));
    CSFM_CUDA(cudaGetLastError());
    idx->stats.kernel_launches += 1;
  }
  if (a.steps_total) {  // ctr[1] = rank steps, ctr[2] = k-mer table lookups
    CSFM_CUDA(cudaMemcpyAsync((unsigned long long*)idx->h_pinned + 8, ctr + 1, 8, cudaMemcpyDeviceToHost, stream));
    CSFM_CUDA(cudaMemcpyAsync((unsigned long long*)idx->h_pinned + 10, ctr + 2, 32, cudaMemcpyDeviceToHost, stream));  // + ctr[4] = half steps, ctr[5] = level lines fetched
  }
  return CSFM_OK;
}

// Pass 1+2 of locate: intervals, then out_offs = exclusive prefix of min(count, limit).
// Synchronises `stream` (the total decides the size of the position buffer).
int locate_plan(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat, uint64_t limit,
                uint64_t* d_out_offs, int32_t* d_status, uint64_t* total, cudaStream_t stream) {
  *total = 0;
  if (npat == 0) {
    CSFM_CUDA(cudaMemsetAsync(d_out_offs, 0, 8, stream));
    CSFM_CUDA(cudaStreamSynchronize(stream));
    return CSFM_OK;
  }
  // a device-pointer caller may run locate on several streams: the previous call's expand kernel must be done
  // with the intervals in ws_tmp before this call overwrites them
  if (idx->ws_tmp_busy) CSFM_CUDA(cudaStreamWaitEvent(stream, idx->ev_ws_tmp, 0));
  int rc = idx->ws_tmp.ensure((npat + 1) * 8 + 256);
  if (rc) return rc;
  uint32_t* d_row_sp = idx->ws_tmp.as<uint32_t>();
  uint32_t* d_row_cnt = d_row_sp + (npat + 1);
  const uint32_t saved_mask = idx->instr_mask;
  idx->instr_mask &= ~3u;  // in locate the instrumentation belongs to the walk kernel
  rc = count_device(idx, d_bytes, d_offs, npat, nullptr, nullptr, d_row_sp, d_row_cnt, limit, stream);
  idx->instr_mask = saved_mask;
  if (rc) return rc;
  // npat + 1 items with a zero tail: out_offs[npat] = the total comes out of the same scan, on the device
  CSFM_CUDA(cudaMemsetAsync(d_row_cnt + npat, 0, 4, stream));
  // u32 inputs accumulated into u64 outputs (the init value's type drives the accumulator)
  size_t tmp_bytes = 0;
  CSFM_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, d_row_cnt, d_out_offs, cub::Sum(), (uint64_t)0,
                                           (int64_t)(npat + 1), stream));
  rc = idx->ws_scan.ensure(tmp_bytes + 16);
  if (rc) return rc;
  CSFM_CUDA(cub::DeviceScan::ExclusiveScan(idx->ws_scan.p, tmp_bytes, d_row_cnt, d_out_offs, cub::Sum(), (uint64_t)0,
                                           (int64_t)(npat + 1), stream));
  idx->stats.kernel_launches += 2;  // cub scan: init + scan kernels
  if (d_status) CSFM_CUDA(cudaMemsetAsync(d_status, 0, npat * 4, stream));
  uint64_t* h = reinterpret_cast<uint64_t*>(idx->h_pinned) + 16 + (idx->total_slot++ & 15u);  // a slot of its own per call
  CSFM_CUDA(cudaMemcpyAsync(h, d_out_offs + npat, 8, cudaMemcpyDeviceToHost, stream));
  CSFM_CUDA(cudaStreamSynchronize(stream));
  *total = *h;
  return CSFM_OK;
}

// Pass 3+4 of locate: SA rows of every slot, then the LF walks. Asynchronous on `stream`.
// Uses the intervals left in idx->ws_tmp by locate_plan.
// Index with the full suffix array on board (text-verification sections: the text ends in a unique
// smallest byte, so the LF walk from row r ends at (sample + steps) % n == SA[r] and never fails):
// the position of a row is one gather instead of ~stride LF steps.
__global__ void rows_to_positions_kernel(const uint32_t* __restrict__ sa, uint64_t* __restrict__ out_pos,
                                         unsigned long long total, uint32_t row_base, uint32_t rows_implicit) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x)
    out_pos[i] = sa[rows_implicit ? row_base + (uint32_t)i : (uint32_t)out_pos[i]];
}

int locate_expand(csfm_index* idx, uint64_t npat, const uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t total,
                  cudaStream_t stream) {
  if (npat == 0 || total == 0) return CSFM_OK;
  const uint32_t* d_row_sp = idx->ws_tmp.as<uint32_t>();
  expand_rows_kernel<<<idx->num_sms * 8, 256, 0, stream>>>(d_row_sp, d_out_offs, npat, d_out_pos);
  CSFM_CUDA(cudaGetLastError());
  if (!idx->ev_ws_tmp) CSFM_CUDA(cudaEventCreateWithFlags(&idx->ev_ws_tmp, cudaEventDisableTiming));
  CSFM_CUDA(cudaEventRecord(idx->ev_ws_tmp, stream));  // ws_tmp (the intervals) is free again from here on
  idx->ws_tmp_busy = true;
  idx->stats.kernel_launches += 1;
  return CSFM_OK;
}

int locate_walk(csfm_index* idx, uint64_t npat, const uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t first,
                uint64_t count, int32_t* d_status, cudaStream_t stream, int64_t row_base) {
  if (npat == 0 || count == 0) return CSFM_OK;
  const uint32_t rows_implicit = row_base >= 0 ? 1u : 0u;
  if (idx->view.dense && idx->view.dense_shift == 0 && !idx->no_sa_locate) {
    const bool timed_sa = (idx->instr_mask & 2u) != 0;
    if (timed_sa) CSFM_CUDA(cudaEventRecord(idx->ev0, stream));
    const int g = (int)std::min<uint64_t>((count + 255) / 256, (uint64_t)idx->num_sms * 8);
    rows_to_positions_kernel<<<g, 256, 0, stream>>>(idx->view.dense, d_out_pos + first, count, (uint32_t)std::max<int64_t>(row_base, 0) + (uint32_t)first,
                                                    rows_implicit);
    if (timed_sa) CSFM_CUDA(cudaEventRecord(idx->ev1, stream));
    CSFM_CUDA(cudaGetLastError());
    idx->stats.kernel_launches += 1;
    return CSFM_OK;
  }
  unsigned long long* ctr = next_counter_slot(idx);
  CSFM_CUDA(cudaMemsetAsync(ctr, 0, kCounterWords * sizeof(unsigned long long), stream));
  WalkArgs w{};
  w.row_base = (uint32_t)std::max<int64_t>(row_base, 0);
  w.rows_implicit = rows_implicit;
  w.out_pos = d_out_pos;
  w.first = first;
  w.total = count;
  w.out_offs = d_out_offs;
  w.npat = npat;
  w.status = d_status;
  w.cursor = ctr;
  w.lf_total = (idx->instr_mask & 1u) ? ctr + 1 : nullptr;
  const bool nib = idx->view.layout == kLayoutNibble128;
  const bool dna = idx->view.layout == kLayoutDna64;
  // layout 3: one lane per row while the walk's working set (level lines + samples) is within reach of the L2 — measured
  // on C4: 3.3e9 against 3.05e9 occ/s with row samples (123 MB), 5.5e9 against 4.1e9 with position samples (168 MB) — a
  // two-lane sub-warp (one 64-byte request per line) when the lines come out of HBM (tools/ab_walk3_lanes.sh,
  // tools/ab_lanes_ctas.sh, tools/ab_position_samples.sh)
  const uint64_t walk_set = (uint64_t)idx->view.L * idx->view.level_stride + (uint64_t)idx->view.nsamp * 4;
  const int lanes3 = idx->walk3_lanes ? idx->walk3_lanes : (walk_set <= (256ull << 20) ? 1 : 2);
  const int grid_max = dna ? idx->num_sms * max_blocks_per_sm_walk3(idx->view, lanes3)
                           : nib ? idx->num_sms * max_blocks_per_sm_walk2() : persistent_grid(idx, (const void*)walk_kernel);
  const uint64_t want = (count * (dna ? lanes3 : 4) + kThreads - 1) / kThreads;
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)grid_max);
  const bool timed = (idx->instr_mask & 2u) != 0;
  if (timed) CSFM_CUDA(cudaEventRecord(idx->ev0, stream));
  if (dna)
    launch_walk3(idx->view, w, grid, stream, lanes3);
  else if (nib)
    launch_walk2(idx->view, w, grid, stream);
  else
    walk_kernel<<<grid, kThreads, 0, stream>>>(idx->view, w);
  if (timed) CSFM_CUDA(cudaEventRecord(idx->ev1, stream));
  CSFM_CUDA(cudaGetLastError());
  idx->stats.kernel_launches += 1;
  if (w.lf_total)
    CSFM_CUDA(cudaMemcpyAsync((unsigned long long*)idx->h_pinned + 9, ctr + 1, 8, cudaMemcpyDeviceToHost, stream));
  return CSFM_OK;
}

int offsets_from_lengths8(const uint8_t* d_lens, uint64_t count, uint64_t* d_offs, DeviceBuffer& scratch,
                          cudaStream_t stream) {
  // u8 inputs accumulated into u64 outputs (the init value's type drives the accumulator)
  size_t tmp_bytes = 0;
  CSFM_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, d_lens, d_offs, cub::Sum(), (uint64_t)0, (int64_t)count, stream));
  int rc = scratch.ensure(tmp_bytes + 16);
  if (rc) return rc;
  CSFM_CUDA(cub::DeviceScan::ExclusiveScan(scratch.p, tmp_bytes, d_lens, d_offs, cub::Sum(), (uint64_t)0, (int64_t)count, stream));
  return CSFM_OK;
}

int extract_bwt_device(csfm_index* idx, uint8_t* d_out, cudaStream_t stream) {
  if (idx->h.n == 0) return CSFM_OK;
  const bool nib = idx->view.layout == kLayoutNibble128;
  const bool dna = idx->view.layout == kLayoutDna64;
  int per_sm = 0;
  if (dna)
    per_sm = max_blocks_per_sm_access3(idx->view);
  else if (nib)
    per_sm = max_blocks_per_sm_access2();
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, access_kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const uint64_t want = (idx->h.n * 4 + kThreads - 1) / kThreads;
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)idx->num_sms * per_sm);
  if (dna)
    launch_access3(idx->view, d_out, grid, stream);
  else if (nib)
    launch_access2(idx->view, d_out, grid, stream);
  else
    access_kernel<<<grid, kThreads, 0, stream>>>(idx->view, d_out);
  CSFM_CUDA(cudaGetLastError());
  return CSFM_OK;
}

}  // namespace csfm
