// csfm_kernels.cuh — pieces shared by the query kernels of both layouts: the per-CTA table
// copy, the warp-local work queue and the kernel argument blocks.
#pragma once
#include "csfm_common.cuh"

namespace csfm {

constexpr int kThreads = 256;
constexpr uint32_t kChunk = 32;  // queries fetched per warp per global atomic

struct Tables {  // per-CTA shared-memory copy of the byte-indexed tables
  uint32_t C[257];
  uint32_t pad[3];
  uint32_t base_by_byte[256];
  uint32_t base_by_code[256];
  uint32_t start1[16];
  uint8_t code_of_byte[256];
  uint8_t byte_of_code[256];
};

__device__ __forceinline__ void load_tables(Tables& t, const BlobHeader* __restrict__ h) {
  for (int i = threadIdx.x; i < 257; i += blockDim.x) t.C[i] = h->C[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    t.base_by_byte[i] = h->base_by_byte[i];
    t.base_by_code[i] = h->base_by_code[i];
    t.code_of_byte[i] = h->code_of_byte[i];
    t.byte_of_code[i] = h->byte_of_code[i];
  }
  if (threadIdx.x < 16) t.start1[threadIdx.x] = h->start1[threadIdx.x];
  __syncthreads();
}

// Everything a backward-search step needs to know about its byte, in one 16-byte shared load:
// x = base_by_byte, y = compact code (bit 31 set: byte absent from the text), z = start1[code >> 4].
// Call before load_tables() (which ends with the CTA barrier).
__device__ __forceinline__ void load_step_table(uint4* step, const BlobHeader* __restrict__ h) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const uint32_t code = h->code_of_byte[i];
    const bool absent = h->C[i + 1] == h->C[i];
    step[i] = make_uint4(h->base_by_byte[i], code | (absent ? 0x80000000u : 0u), h->start1[code >> 4], 0u);
  }
}

// Warp-local work queue over [0,total): returns this sub-warp's next item or ~0ull.
struct WarpQueue {
  unsigned long long next = 0, end = 0;
  bool exhausted = false;
};

// Called by all 32 lanes (converged). `need` = this sub-warp wants an item. Returns the item
// index or ~0ull. A partially served round simply leaves some sub-warps idle for one trip.
// min_idle > 1 batches the refills of a warp: nothing is handed out until that many sub-warps are
// idle, so the refill code (issued for the whole warp whoever needs it) runs on fewer trips.
// `served` (warp-uniform) tells whether anything was handed out.
__device__ __forceinline__ unsigned long long queue_take(WarpQueue& q, bool need, int lane,
                                                         unsigned long long* cursor,
                                                         unsigned long long total, unsigned min_idle, bool& served) {
  served = false;
  const unsigned need_mask = __ballot_sync(0xFFFFFFFFu, need) & 0x11111111u;  // leaders
  if ((unsigned)__popc(need_mask) < min_idle) return ~0ull;
  if (q.next >= q.end && !q.exhausted) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= total) {
      q.exhausted = true;
    } else {
      q.next = base;
      q.end = (base + kChunk < total) ? base + kChunk : total;
    }
  }
  const unsigned long long avail = q.end - q.next;
  const unsigned my_rank = __popc(need_mask & ((1u << (lane & ~3)) - 1u));
  const unsigned cnt = __popc(need_mask);
  unsigned long long item = ~0ull;
  if (need && my_rank < avail) item = q.next + my_rank;
  q.next += (cnt < avail) ? cnt : avail;
  served = avail != 0;
  return item;
}
__device__ __forceinline__ unsigned long long queue_take(WarpQueue& q, bool need, int lane,
                                                         unsigned long long* cursor, unsigned long long total) {
  bool served;
  return queue_take(q, need, lane, cursor, total, 1u, served);
}

// 32-bit form for kernels whose launches stay below 2^32 items (count_device slices at 2^31):
// two registers less per thread. Returns the item index or ~0u.
struct WarpQueue32 {
  uint32_t next = 0, end = 0;
  bool exhausted = false;
};
__device__ __forceinline__ uint32_t queue_take32(WarpQueue32& q, bool need, int lane, unsigned long long* cursor,
                                                 uint32_t total, unsigned min_idle, bool& served) {
  served = false;
  const unsigned need_mask = __ballot_sync(0xFFFFFFFFu, need) & 0x11111111u;  // leaders
  if ((unsigned)__popc(need_mask) < min_idle) return ~0u;
  if (q.next >= q.end && !q.exhausted) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= total) {
      q.exhausted = true;
    } else {
      q.next = (uint32_t)base;
      q.end = (total - q.next > kChunk) ? q.next + kChunk : total;
    }
  }
  const uint32_t avail = q.end - q.next;
  const unsigned my_rank = __popc(need_mask & ((1u << (lane & ~3)) - 1u));
  const unsigned cnt = __popc(need_mask);
  uint32_t item = ~0u;
  if (need && my_rank < avail) item = q.next + my_rank;
  q.next += (cnt < avail) ? cnt : avail;
  served = avail != 0;
  return item;
}

// The same two queues for sub-warps of G lanes (G = 2: layout 3, sixteen queries per warp). The leader of a
// sub-warp is its lowest lane; `need` must be uniform within a sub-warp.
template <int G>
__device__ __forceinline__ constexpr unsigned leader_mask() { return G == 1 ? 0xFFFFFFFFu : G == 2 ? 0x55555555u : G == 4 ? 0x11111111u : 0x01010101u; }

template <int G>
__device__ __forceinline__ uint32_t queue_take32g(WarpQueue32& q, bool need, int lane, unsigned long long* cursor, uint32_t total) {
  const unsigned need_mask = __ballot_sync(0xFFFFFFFFu, need) & leader_mask<G>();
  if (need_mask == 0) return ~0u;
  if (q.next >= q.end && !q.exhausted) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= total) {
      q.exhausted = true;
    } else {
      q.next = (uint32_t)base;
      q.end = (total - q.next > kChunk) ? q.next + kChunk : total;
    }
  }
  const uint32_t avail = q.end - q.next;
  const unsigned my_rank = __popc(need_mask & ((1u << (lane & ~(G - 1))) - 1u));
  const unsigned cnt = __popc(need_mask);
  uint32_t item = ~0u;
  if (need && my_rank < avail) item = q.next + my_rank;
  q.next += (cnt < avail) ? cnt : avail;
  return item;
}

template <int G>
__device__ __forceinline__ unsigned long long queue_takeg(WarpQueue& q, bool need, int lane, unsigned long long* cursor,
                                                          unsigned long long total) {
  const unsigned need_mask = __ballot_sync(0xFFFFFFFFu, need) & leader_mask<G>();
  if (need_mask == 0) return ~0ull;
  if (q.next >= q.end && !q.exhausted) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= total) {
      q.exhausted = true;
    } else {
      q.next = base;
      q.end = (base + kChunk < total) ? base + kChunk : total;
    }
  }
  const unsigned long long avail = q.end - q.next;
  const unsigned my_rank = __popc(need_mask & ((1u << (lane & ~(G - 1))) - 1u));
  const unsigned cnt = __popc(need_mask);
  unsigned long long item = ~0ull;
  if (need && my_rank < avail) item = q.next + my_rank;
  q.next += (cnt < avail) ? cnt : avail;
  return item;
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX): pattern staging ------------------------
// cp.async.bulk moves a contiguous, 16-byte aligned span global -> shared through the TMA unit
// (SASS: UBLKCP) and signals an mbarrier with the byte count; nobody spends registers or issue
// slots on the copy.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a copy that never lands is a bug — trap (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that occupy no registers -----------
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Text-verification shortcut: once the interval of a query is down to at most kVerifyRows rows
// and at most kVerifyMax characters are left, lane j of the sub-warp compares those characters
// with the text in front of the suffix of row sp + j. The windows (one text window per row, one
// pattern window) are staged in a per-sub-warp slot as aligned 16-byte chunks: a window of 32
// unaligned bytes spans at most three of them.
constexpr uint32_t kVerifyMax = 32;
constexpr uint32_t kVerifyRows = 4;
struct alignas(16) VerifySlot {
  uint8_t t[kVerifyRows][48];
  uint8_t p[48];
  uint8_t pad[16];
};

// t[toff + k] == p[poff + k] for all k < rem (rem <= kVerifyMax, toff and poff < 16), four
// characters at a time: both windows are re-aligned with funnel shifts, the bytes past rem masked.
__device__ __forceinline__ bool windows_equal(const uint8_t* t, const uint8_t* p, uint32_t toff, uint32_t poff,
                                              uint32_t rem) {
  const uint32_t* tw = reinterpret_cast<const uint32_t*>(t) + (toff >> 2);
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(p) + (poff >> 2);
  const uint32_t ts = (toff & 3u) * 8u, ps = (poff & 3u) * 8u;
  uint32_t acc = 0, tlo = tw[0], plo = pw[0];
#pragma unroll
  for (uint32_t w = 0; w < kVerifyMax / 4; ++w) {
    const uint32_t thi = tw[w + 1], phi = pw[w + 1];
    const uint32_t x = __funnelshift_r(tlo, thi, ts) ^ __funnelshift_r(plo, phi, ps);
    const int left = (int)rem - (int)(4 * w);                 // characters from this word on
    const uint32_t valid = (uint32_t)min(max(left, 0), 4);    // ... that belong to this word
    acc |= x & __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * valid);
    tlo = thi;
    plo = phi;
  }
  return acc == 0;
}

constexpr uint32_t kStageBytes = 2048;  // pattern bytes staged per warp chunk (32 patterns)
constexpr uint32_t kPrivBytes = 64;     // private slot per sub-warp for the pattern it is working on

// Per-warp double buffer: the packed bytes of a 32-pattern chunk and its kChunk+1 offsets.
struct alignas(16) WarpStage {
  uint8_t bytes[2][kStageBytes];
  uint64_t offs[2][kChunk + 2];
  const uint8_t* base_ptr[2];  // pattern byte at batch offset o lives at base_ptr[b] + o (shared or global)
  uint64_t bar_offs[2];
  uint64_t bar_bytes[2];
};

struct CountArgs {
  const uint8_t* bytes;
  const uint64_t* offs;
  unsigned long long npat;
  uint64_t* counts;    // nullable
  uint64_t* sp_ep;     // nullable, 2 per query
  uint32_t* row_sp;    // nullable (locate pass 1)
  uint32_t* row_cnt;   // nullable (locate pass 1): min(count, limit), 0 for empty patterns
  uint32_t limit32;
  unsigned long long* cursor;
  unsigned long long* steps_total;  // nullable (instrumentation)
  // Two-pass count (count2q_kernel, then the sub-warp kernel over what it could not finish): the first pass
  // appends the queries it leaves to qlist and counts them in *qlist_len; the second pass takes its items
  // from that list instead of [0, npat). Both null for a one-pass launch.
  uint32_t* qlist;
  unsigned long long* qlist_len;
};

// Single-query path (cs::FMIndex::count / locate called one pattern at a time, tools/benchmark.cpp): the
// pattern travels as a kernel PARAMETER and the result comes back through mapped pinned memory that
// the host spins on, so a query is one launch: no staging copies, no stream synchronisation.
constexpr uint32_t kSingleMax = 64;  // longer single patterns go through the batch path
struct SingleQuery {
  // what a step needs to know about each pattern character, looked up ON THE HOST from its copy of the header tables
  // and handed over as kernel parameters (constant bank): x = base, y = compact code | absent << 31, z = start1[hi], w = C[byte]
  uint4 ent[kSingleMax];
  uint32_t len;
  uint32_t c_after_last;  // C[last byte + 1]: the end of the interval of a pattern that starts from the C array
  uint32_t seq;           // stored with the result: the host waits for it
  uint32_t pad;
};
struct alignas(16) SingleResult {  // ONE 16-byte store: count, interval and the sequence number land together
  unsigned int count, sp, ep, seq;   // (counts and rows are < 2^32: n < 2^32 - 1)
};

struct WalkArgs {
  uint32_t row_base;       // rows_implicit: slot i starts at SA row row_base + i (single-query locate) ...
  uint32_t rows_implicit;  // ... instead of reading its row from out_pos[i]
  uint64_t* out_pos;  // in: SA row, out: text position
  unsigned long long first;  // this launch walks the output slots [first, first + total)
  unsigned long long total;
  const uint64_t* out_offs;  // npat+1 (to attribute a failed walk to its query)
  unsigned long long npat;
  int32_t* status;
  unsigned long long* cursor;
  unsigned long long* lf_total;  // nullable
};

// Layout-2 kernels (csfm_query2.cu); launched by the dispatchers in csfm_query.cu.
void launch_count2(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream, bool tma_staging);
// One query per THREAD for the queries that finish in "table lookup, (half step,) text verification"; the rest go to a.qlist.
bool count2q_eligible(const IndexView& iv, const CountArgs& a);
void launch_count2q(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream);
int max_blocks_per_sm_count2q(const CountArgs& a);
void launch_walk2(const IndexView& iv, const WalkArgs& a, int grid, cudaStream_t stream);
void launch_count_single2(const IndexView& iv, const SingleQuery& q, SingleResult* d_result, cudaStream_t stream);
void launch_access2(const IndexView& iv, uint8_t* out, int grid, cudaStream_t stream);
int max_blocks_per_sm_count2(bool tma_staging, const IndexView& iv, const CountArgs& a);
int max_blocks_per_sm_walk2();
int max_blocks_per_sm_access2();
// Layout-3 kernels (csfm_query3.cu)
void launch_count3(const IndexView& iv, const CountArgs& a, int grid, cudaStream_t stream, int lanes);  // lanes per query: 2, or 1 for L2-resident indexes
void launch_walk3(const IndexView& iv, const WalkArgs& a, int grid, cudaStream_t stream, int lanes);
void launch_access3(const IndexView& iv, uint8_t* out, int grid, cudaStream_t stream);
int max_blocks_per_sm_count3(const IndexView& iv, const CountArgs& a, int lanes);
// the text back out of the index (extract without a text section): every sampled row walks LF to the next one
void launch_untext2(const IndexView& iv, uint8_t* out, unsigned long long* cursor, unsigned long long* written, int num_sms, cudaStream_t stream);
void launch_untext3(const IndexView& iv, uint8_t* out, unsigned long long* cursor, unsigned long long* written, int num_sms, cudaStream_t stream);
int max_blocks_per_sm_walk3(const IndexView& iv, int lanes);
int max_blocks_per_sm_access3(const IndexView& iv);

}  // namespace csfm
