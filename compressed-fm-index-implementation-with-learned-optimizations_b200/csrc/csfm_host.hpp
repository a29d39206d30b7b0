// csfm_host.hpp — host-side handle and helpers shared by the library's translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include <cuda_runtime.h>

#include "../../include/csfm.h"
#include "csfm_common.cuh"

namespace csfm {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define CSFM_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return ::csfm::fail(_e == cudaErrorMemoryAllocation ? CSFM_ERR_NOMEM : CSFM_ERR_CUDA, \
                          std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    }                                                                                       \
  } while (0)

// CSFM_BUILD_TIMERS=1: wall-clock of the construction phases on stderr (each mark synchronises the device).
struct PhaseTimer {
  bool on;
  double t0;
  const char* what;
  explicit PhaseTimer(const char* w);
  void mark(const char* phase);
};

// Grow-only device buffer (workspace reuse across batch calls).
struct DeviceBuffer {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);  // returns csfm_status
  void release();
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};

}  // namespace csfm

struct csfm_index {
  int device = 0;
  int num_sms = 0;
  uint8_t* d_blob = nullptr;
  uint64_t blob_bytes = 0;
  bool owns_blob = true;
  csfm::BlobHeader h{};  // host copy of the header
  csfm::IndexView view{};
  uint32_t* d_sa = nullptr;  // CSFM_BUILD_KEEP_SA
  uint32_t sa_rounds = 0, sa_radix_passes = 0;  // how the suffix sort of build_from_text went (0: index not built here)
  uint64_t sa_pair_passes = 0;

  cudaStream_t stream = nullptr;  // used by the host-pointer API
  cudaStream_t aux_stream[2] = {nullptr, nullptr};  // slice pipeline of large host-pointer batches
  struct AsyncSlot {  // csfm_count_batch_submit / _wait
    cudaStream_t stream = nullptr;
    csfm::DeviceBuffer in, out, scan;
    uint64_t ticket = 0;  // ticket currently occupying the slot (0 = free)
  } async_slot[CSFM_ASYNC_SLOTS];
  uint64_t next_ticket = 1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_ws_tmp = nullptr;  // locate: recorded after the last reader of ws_tmp (expand_rows)
  bool ws_tmp_busy = false;
  uint32_t total_slot = 0;  // locate: rotating pinned slot the total is read back into
  cudaEvent_t ev_slice[4] = {nullptr, nullptr, nullptr, nullptr};  // locate: walk slice done -> its copy may start
  csfm::DeviceBuffer ws_in, ws_out, ws_tmp, ws_scan, ws_pos;
  unsigned long long* d_counters = nullptr;  // ring of work cursors / accumulators
  uint32_t counter_slot = 0;
  void* h_pinned = nullptr;  // small pinned scratch (totals, stats), mapped: the single-query kernel writes its result here
  void* d_pinned = nullptr;  // device alias of h_pinned
  uint32_t single_seq = 0;   // sequence number of the last single-query launch
  struct SingleCache {       // interval of the last single-pattern locate: the sizing call and the fill call search once
    bool valid = false;
    uint32_t len = 0;
    uint8_t bytes[256];
    uint64_t sp = 0, count = 0;
  } single_cache;

  // two-pass count (count2q_kernel + the sub-warp kernel over what it left): a small ring of overflow lists, each
  // guarded by an event so that a list is not reused while an earlier launch may still read it
  struct QListSlot {
    csfm::DeviceBuffer buf;
    cudaEvent_t done = nullptr;
    uint64_t npat = 0;  // batch size of the launch that used the slot last (0 = never used)
  } qslot[8];
  uint32_t qslot_next = 0;
  uint32_t two_pass_skip = 0;   // calls for which the two-pass form is skipped (most queries of recent batches overflowed)
  int count3_lanes = 0;         // CSFM_COUNT3_LANES=1|2: lanes per query of the layout-3 count kernel (0: by index size)
  int walk3_lanes = 0;          // CSFM_WALK3_LANES=1|2: lanes per row of the layout-3 walk kernel (0: by index size)
  bool no_two_pass = true;      // the two-pass form is opt-in (CSFM_TWO_PASS=1): measured slower than the one-pass sub-warp kernel

  uint8_t* d_text_cache = nullptr;  // csfm_extract on an index without a text section: the text, rebuilt once by LF walks

  uint32_t instr_mask = 0;
  bool tma_staging = false;  // count kernel variant (csfm_set_option / CSFM_PATTERN_STAGING=tma)
  bool no_sa_locate = false;  // CSFM_NO_SA_LOCATE: walk even when the index carries its suffix array
  csfm_call_stats stats{};
  std::mutex mu;
};

namespace csfm {

constexpr uint32_t kCounterSlots = 256;
constexpr uint32_t kCounterWords = 8;  // u64 per slot: work cursor, rank steps, table lookups, text checks, half steps

// csfm_build.cu
// d_text / d_sa (both nullable): when given and the text ends in a unique smallest byte, they are
// copied into the blob for the verification shortcut.
int index_from_device_bwt(const uint8_t* d_bwt, uint64_t n, const uint32_t* d_ssa, uint64_t nsamp,
                          uint32_t stride, int device, uint32_t flags, csfm_index** out,
                          const uint8_t* d_text = nullptr, const uint32_t* d_sa = nullptr);
int index_finish_handle(csfm_index* idx);  // fills view/stream/workspace after d_blob + h are set
int build_kmer_table(csfm_index* idx, cudaStream_t stream);  // csfm_query2.cu: fills the table sections
int build_kmer_table3(csfm_index* idx, cudaStream_t stream);  // csfm_query3.cu: the same for layout 3
// csfm_sa.cu
int build_sa_bwt_device(const uint8_t* d_text, uint64_t n, uint32_t stride, cudaStream_t stream,
                        uint8_t** d_bwt_out, uint32_t** d_ssa_out, uint64_t* nsamp_out,
                        uint32_t** d_sa_out /*nullable: keep SA*/, uint32_t* rounds_out = nullptr,
                        uint32_t* passes_out = nullptr, uint64_t* pair_passes_out = nullptr);
// csfm_query.cu
int count_device(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat,
                 uint64_t* d_counts, uint64_t* d_sp_ep, uint32_t* d_row_sp, uint32_t* d_row_cnt,
                 uint64_t limit, cudaStream_t stream);
int locate_plan(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat,
                uint64_t limit, uint64_t* d_out_offs, int32_t* d_status, uint64_t* total,
                cudaStream_t stream);
// rows of every output slot (uses the intervals locate_plan left in the workspace) ...
int locate_expand(csfm_index* idx, uint64_t npat, const uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t total,
                  cudaStream_t stream);
// ... then rows -> text positions for the slots [first, first + count): LF walks, or one gather per row
// when the index carries its whole suffix array
// row_base >= 0: slot i starts at SA row row_base + i (one query's interval) instead of the row stored in d_out_pos[i]
int locate_walk(csfm_index* idx, uint64_t npat, const uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t first,
                uint64_t count, int32_t* d_status, cudaStream_t stream, int64_t row_base = -1);
int extract_bwt_device(csfm_index* idx, uint8_t* d_out, cudaStream_t stream);
// d_offs[0..count) = exclusive prefix sum of d_lens[0..count) (u8 lengths -> u64 offsets), on `stream`
int offsets_from_lengths8(const uint8_t* d_lens, uint64_t count, uint64_t* d_offs, DeviceBuffer& scratch,
                          cudaStream_t stream);
unsigned long long* next_counter_slot(csfm_index* idx);

}  // namespace csfm
