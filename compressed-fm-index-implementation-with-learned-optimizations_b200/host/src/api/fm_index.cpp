// host/src/api/fm_index.cpp — cs::FMIndex over the C ABI of libcsfm.so (include/csfm.h).
// Mirrors the behaviour of /root/reference/src/api/fm_index.cpp; contains no FM-index algorithm.
#include "fm_index.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>

#include "../../../../include/csfm.h"
#include "../serialization/csidx.hpp"
#include "../util/timer.hpp"

namespace cs {

namespace {

int g_default_device = -1;

int default_device() {
  if (g_default_device >= 0) return g_default_device;
  if (const char* e = std::getenv("CS_DEVICE")) return std::atoi(e);
  return 0;
}

[[noreturn]] void throw_last(const char* what) {
  throw std::runtime_error(std::string(what) + ": " + csfm_last_error());
}

void pack(const std::vector<std::string_view>& pats, std::vector<uint8_t>& bytes, std::vector<uint64_t>& offs) {
  offs.resize(pats.size() + 1);
  uint64_t total = 0;
  for (size_t i = 0; i < pats.size(); ++i) {
    offs[i] = total;
    total += pats[i].size();
  }
  offs[pats.size()] = total;
  bytes.resize(total ? total : 1);
  for (size_t i = 0; i < pats.size(); ++i)
    if (!pats[i].empty()) std::memcpy(bytes.data() + offs[i], pats[i].data(), pats[i].size());
}

}  // namespace

void FMIndex::set_default_device(int device) { g_default_device = device; }

csfm_index* FMIndex::local() const {
  if (!handle_ || std::this_thread::get_id() == owner_) return handle_.get();
  // a few aliases per thread, most recently used first; an entry keeps the index that owns the blob alive
  struct Entry {
    std::shared_ptr<csfm_index> owner, alias;
  };
  static thread_local std::vector<Entry> cache;
  for (size_t i = 0; i < cache.size(); ++i)
    if (cache[i].owner.get() == handle_.get()) {
      if (i) std::swap(cache[i], cache[0]);
      return cache[0].alias.get();
    }
  csfm_index* a = nullptr;
  if (csfm_alias(handle_.get(), &a) != CSFM_OK) return handle_.get();  // still correct, only serialised
  if (cache.size() >= 4) cache.pop_back();
  cache.insert(cache.begin(), Entry{handle_, std::shared_ptr<csfm_index>(a, [](csfm_index* x) { csfm_destroy(x); })});
  return a;
}

FMIndex FMIndex::build_from_text(const std::string& text, const BuildParams& p) {
  return build_from_text(text, p, CSFM_BUILD_DEFAULT);
}

FMIndex FMIndex::build_from_text(const std::string& text, const BuildParams& p, uint32_t csfm_build_flags) {
  FMIndex idx;
  idx.meta_.n = text.size();
  idx.text_ = std::make_shared<const std::string>(text);
  csfm_params cp{p.S, p.s, p.ssa_stride, p.eps};
  csfm_index* h = nullptr;
  {
    // The reference prints four [TIMER] lines here (fm_index.cpp:26,31,50,57); the whole build is
    // one device pipeline now, reported as a single line when CS_TIMERS=1.
    const bool timers = std::getenv("CS_TIMERS") != nullptr;
    std::unique_ptr<ScopeTimer> t(timers ? new ScopeTimer("build_from_text(gpu)") : nullptr);
    if (csfm_build_from_text(reinterpret_cast<const uint8_t*>(text.data()), text.size(), &cp, default_device(),
                             csfm_build_flags, &h) != CSFM_OK)
      throw_last("build_from_text");
  }
  idx.handle_ = std::shared_ptr<csfm_index>(h, [](csfm_index* x) { csfm_destroy(x); });
  return idx;
}

FMIndex FMIndex::open_directory(const std::string&) {
  throw std::runtime_error("on-disk open not implemented yet");  // fm_index.cpp:72
}

int FMIndex::device() const {
  csfm_index_info info;
  if (!handle_ || csfm_info(handle_.get(), &info) != CSFM_OK) return -1;
  return static_cast<int>(info.device);
}

FMIndex FMIndex::replicate_to(int device) const {
  if (!handle_) throw std::runtime_error("replicate_to: index not built");
  csfm_index* h = nullptr;
  if (csfm_replicate(handle_.get(), device, &h) != CSFM_OK) throw_last("replicate_to");
  FMIndex r;
  r.meta_ = meta_;
  r.text_ = text_;
  r.handle_ = std::shared_ptr<csfm_index>(h, [](csfm_index* x) { csfm_destroy(x); });
  return r;
}

void FMIndex::count_batch_sharded(const std::vector<FMIndex>& replicas, const uint8_t* bytes, const uint64_t* offs,
                                  uint64_t npat, uint64_t* counts) {
  if (replicas.empty()) throw std::runtime_error("count_batch_sharded: no replicas");
  const uint64_t world = replicas.size();
  std::vector<std::thread> workers;
  std::vector<std::exception_ptr> errors(world);
  for (uint64_t r = 0; r < world; ++r) {
    const uint64_t lo = npat * r / world, hi = npat * (r + 1) / world;
    if (hi == lo) continue;
    workers.emplace_back([&, r, lo, hi] {
      try {
        // a slice keeps the batch's absolute offsets: rebase them so that the slice is a batch of its own
        std::vector<uint64_t> local(hi - lo + 1);
        for (uint64_t i = lo; i <= hi; ++i) local[i - lo] = offs[i] - offs[lo];
        // every replica is driven by exactly one worker here: its own handle, no per-thread alias needed
        if (csfm_count_batch(replicas[r].handle(), bytes + offs[lo], local.data(), hi - lo, counts + lo, nullptr) != CSFM_OK)
          throw_last("count_batch_sharded");
      } catch (...) {
        errors[r] = std::current_exception();
      }
    });
  }
  for (auto& w : workers) w.join();
  for (auto& e : errors)
    if (e) std::rethrow_exception(e);
}

void FMIndex::count_batch(const uint8_t* bytes, const uint64_t* offs, uint64_t npat, uint64_t* counts,
                          uint64_t* sp_ep) const {
  if (!handle_) throw std::runtime_error("count_batch: index not built");
  if (csfm_count_batch(local(), bytes, offs, npat, counts, sp_ep) != CSFM_OK) throw_last("count_batch");
}

std::vector<uint64_t> FMIndex::count_batch(const std::vector<std::string_view>& patterns) const {
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> offs;
  pack(patterns, bytes, offs);
  std::vector<uint64_t> counts(patterns.size());
  if (!patterns.empty()) count_batch(bytes.data(), offs.data(), patterns.size(), counts.data());
  return counts;
}

std::vector<uint64_t> FMIndex::count_batch(const std::vector<std::string>& patterns) const {
  std::vector<std::string_view> v(patterns.begin(), patterns.end());
  return count_batch(v);
}

uint64_t FMIndex::count(std::string_view pattern) const {
  if (pattern.empty()) return meta_.n;  // fm_index.cpp:80
  if (meta_.n == 0) return 0;           // fm_index.cpp:81
  const uint64_t offs[2] = {0, pattern.size()};
  uint64_t c = 0;
  count_batch(reinterpret_cast<const uint8_t*>(pattern.data()), offs, 1, &c);
  return c;
}

LocateBatch FMIndex::locate_batch(const std::vector<std::string_view>& patterns, size_t limit) const {
  if (!handle_) throw std::runtime_error("locate_batch: index not built");
  LocateBatch r;
  r.offsets.assign(patterns.size() + 1, 0);
  r.status.assign(patterns.size(), 0);
  if (patterns.empty()) return r;
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> offs;
  pack(patterns, bytes, offs);
  // One pass when the guess holds: positions land in a buffer sized from the previous call on this index (a
  // single pattern sizes first: the library remembers its interval, so the second call only walks).
  uint64_t total = 0;
  uint64_t cap = patterns.size() == 1 ? 0 : std::max<uint64_t>(*locate_hint_, 64 * patterns.size());
  r.positions.resize(cap);
  csfm_index* const h = local();
  int rc = csfm_locate_batch(h, bytes.data(), offs.data(), patterns.size(), limit, r.offsets.data(),
                             cap ? r.positions.data() : nullptr, cap, r.status.data(), &total);
  if (rc == CSFM_ERR_CAPACITY || (rc == CSFM_OK && cap == 0 && total != 0)) {
    r.positions.resize(total);
    rc = csfm_locate_batch(h, bytes.data(), offs.data(), patterns.size(), limit, r.offsets.data(),
                           r.positions.data(), total, r.status.data(), &total);
  }
  if (rc != CSFM_OK) throw_last("locate_batch");
  r.positions.resize(total);
  if (patterns.size() > 1) *locate_hint_ = total + total / 8;
  return r;
}

std::vector<uint64_t> FMIndex::locate(std::string_view pattern, size_t limit) const {
  if (pattern.empty() || meta_.n == 0) return {};  // fm_index.cpp:109
  LocateBatch r = locate_batch({pattern}, limit);
  if (r.status[0] == CSFM_Q_LF_WALK_EXCEEDED)
    throw std::runtime_error("locate: LF walk exceeded text length");  // fm_index.cpp:137
  if (r.status[0] == CSFM_Q_SSA_OOB)
    throw std::runtime_error("locate: SSA sample index out of range");  // fm_index.cpp:143
  return std::move(r.positions);
}

void FMIndex::save(const std::string& path, bool include_text) const {
  if (!handle_) throw std::runtime_error("save: index not built");
  csfm_index_info info;
  if (csfm_info(handle_.get(), &info) != CSFM_OK) throw_last("save");
  CsidxSections s;
  s.text_len = meta_.n;
  if (text_ && include_text) {
    s.has_text = true;
    s.text = *text_;
  }
  s.bwt.resize(info.n);
  if (info.n && csfm_extract_bwt(handle_.get(), s.bwt.data()) != CSFM_OK) throw_last("save");
  s.c_array.resize(257);
  if (csfm_get_C(handle_.get(), s.c_array.data()) != CSFM_OK) throw_last("save");
  s.has_ssa = true;
  s.ssa_stride = info.ssa_stride;
  s.ssa.resize(info.nsamp);
  if (info.nsamp && csfm_get_ssa(handle_.get(), s.ssa.data()) != CSFM_OK) throw_last("save");
  s.device_blob.resize(info.blob_bytes);
  if (csfm_blob_to_host(handle_.get(), s.device_blob.data(), s.device_blob.size()) != CSFM_OK) throw_last("save");
  write_csidx(path, s);
}

FMIndex FMIndex::load(const std::string& path) {
  CsidxSections s = read_csidx(path);
  FMIndex idx;
  csfm_index* h = nullptr;
  if (!s.device_blob.empty()) {
    if (csfm_from_host_blob(s.device_blob.data(), s.device_blob.size(), default_device(), &h) != CSFM_OK) throw_last("load");
  } else {
    if (!s.has_ssa || s.ssa_stride == 0) throw std::runtime_error("load: file has neither a device blob nor an SSA section");
    if (csfm_build_from_parts(s.bwt.data(), s.bwt.size(), s.ssa.data(), s.ssa.size(), s.ssa_stride, default_device(),
                              CSFM_BUILD_DEFAULT, &h) != CSFM_OK)
      throw_last("load");
  }
  idx.handle_ = std::shared_ptr<csfm_index>(h, [](csfm_index* x) { csfm_destroy(x); });
  csfm_index_info info;
  if (csfm_info(h, &info) != CSFM_OK) throw_last("load");
  idx.meta_.n = info.n;
  if (s.has_text) idx.text_ = std::make_shared<const std::string>(std::move(s.text));
  return idx;
}

std::string FMIndex::extract(uint64_t p, uint64_t len) const {
  if (!text_ && handle_) {
    // loaded without a TEXT section: the text comes back out of the device index (csfm_extract)
    if (p >= meta_.n) return {};  // fm_index.cpp:164
    len = len < meta_.n - p ? len : meta_.n - p;
    std::string out(len, '\0');
    uint64_t got = 0;
    if (csfm_extract(handle_.get(), p, len, reinterpret_cast<uint8_t*>(out.data()), &got) != CSFM_OK) throw_last("extract");
    out.resize(got);
    return out;
  }
  if (!text_ || p >= text_->size()) return {};  // fm_index.cpp:164
  len = len < text_->size() - p ? len : text_->size() - p;
  return text_->substr(p, len);
}

}  // namespace cs
