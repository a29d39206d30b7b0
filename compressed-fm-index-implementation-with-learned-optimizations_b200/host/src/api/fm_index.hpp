#pragma once
// host/src/api/fm_index.hpp — drop-in for the reference header of the same path
// (/root/reference/src/api/fm_index.hpp:9-68): same namespace, same BuildParams / IndexMeta,
// same public methods with the same observable results, so the reference's callers
// (tools/*.cpp, tests/fm_search_tests.cpp, tests/simple_tests.cpp, tests/debug_fm.cpp) compile
// and run unchanged. Underneath, every query is a CUDA kernel launch through the C ABI in
// include/csfm.h (libcsfm.so); there is no CPU implementation behind this class.
//
// Added on top of the reference surface: count_batch / locate_batch (the batched entry points of
// BASELINE.json's north star) and device selection.
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

struct csfm_index;  // include/csfm.h

namespace cs {

struct BuildParams {  // fm_index.hpp:11-14; only ssa_stride is honoured, as in the reference
  uint32_t S = 512, s = 64, ssa_stride = 32;
  double eps = 1.0;
};
struct IndexMeta { uint64_t n = 0; uint32_t sigma = 256; };  // fm_index.hpp:15

/// Result of FMIndex::locate_batch: positions of query q are
/// positions[offsets[q] .. offsets[q+1]) in SA-row order; status[q] != 0 where the reference's
/// locate() would have thrown (1 = "locate: LF walk exceeded text length").
struct LocateBatch {
  std::vector<uint64_t> offsets;
  std::vector<uint64_t> positions;
  std::vector<int32_t> status;
};

class FMIndex {
public:
  /// fm_index.cpp:16-69 — SA, BWT, C, rank structure and sampled SA, all built on the GPU.
  static FMIndex build_from_text(const std::string& text, const BuildParams& p);
  /// Same, with CSFM_BUILD_* flags of include/csfm.h (e.g. CSFM_BUILD_LARGE_TABLE, CSFM_BUILD_FORCE_TEXT_CHECK).
  static FMIndex build_from_text(const std::string& text, const BuildParams& p, uint32_t csfm_build_flags);
  /// fm_index.cpp:71-73 — throws std::runtime_error("on-disk open not implemented yet").
  static FMIndex open_directory(const std::string& dir);

  /// fm_index.cpp:79-101.
  uint64_t count(std::string_view pattern) const;
  /// fm_index.cpp:107-157 — SA-row order; throws std::runtime_error where the reference does.
  std::vector<uint64_t> locate(std::string_view pattern, size_t limit = 100000) const;
  /// fm_index.cpp:163-167.
  std::string extract(uint64_t pos, uint64_t len) const;

  // ---- batched entry points (new) ---------------------------------------------------------
  std::vector<uint64_t> count_batch(const std::vector<std::string_view>& patterns) const;
  std::vector<uint64_t> count_batch(const std::vector<std::string>& patterns) const;
  /// Packed form: `bytes` holds the patterns back to back, offs[npat+1] their starts.
  void count_batch(const uint8_t* bytes, const uint64_t* offs, uint64_t npat, uint64_t* counts,
                   uint64_t* sp_ep = nullptr) const;
  LocateBatch locate_batch(const std::vector<std::string_view>& patterns, size_t limit = 100000) const;

  // ---- more than one GPU in one process (new) -----------------------------------------------
  /// A replica on another device: the device index is copied there (NVLink peer copy where possible);
  /// the host-side text is shared. One process per GPU with a broadcast is the other way (parallel.py).
  FMIndex replicate_to(int device) const;
  /// Packed count over several replicas: the batch is cut into contiguous slices, one per replica, each
  /// slice counted on its replica's device from its own host thread. Results as from count_batch.
  static void count_batch_sharded(const std::vector<FMIndex>& replicas, const uint8_t* bytes, const uint64_t* offs,
                                  uint64_t npat, uint64_t* counts);

  // ---- persistence (new): the reference's .csidx container (src/serialization/serialization.hpp)
  /// Writes TEXT, BWT, C_ARRAY, SSA and the device-resident index blob (host/src/serialization/csidx.hpp).
  /// include_text = false leaves the TEXT section out: extract() on the loaded index then reads the text back
  /// out of the device index (csfm_extract).
  void save(const std::string& path, bool include_text = true) const;
  /// Loads a file written by save(): one read + one host->device copy, no rebuild. Files without
  /// the device blob are rebuilt from their BWT + SSA sections.
  static FMIndex load(const std::string& path);

  // ---- introspection / placement ------------------------------------------------------------
  uint64_t size() const { return meta_.n; }
  int device() const;
  csfm_index* handle() const { return handle_.get(); }
  /// Device used by build_from_text (default 0, or $CS_DEVICE).
  static void set_default_device(int device);

private:
  /// The handle a query call of the calling thread goes through: the index's own handle on the thread that created
  /// it, a thread-local alias over the same device blob (csfm_alias: own streams and workspaces, no copy) on every
  /// other thread — calls on one C handle are serialised by its mutex, so this is what lets the const query
  /// methods run concurrently like the reference's (fm_index.hpp: no mutable state).
  csfm_index* local() const;
  std::thread::id owner_ = std::this_thread::get_id();
  IndexMeta meta_;
  std::shared_ptr<const std::string> text_;  // host copy for extract(), like FMIndex::text_
  std::shared_ptr<csfm_index> handle_;       // device-resident index; copies share it (read-only)
  std::shared_ptr<uint64_t> locate_hint_ = std::make_shared<uint64_t>(0);  // positions of the previous locate_batch: sizes the next one
};

}  // namespace cs
