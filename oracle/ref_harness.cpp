// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" shim around the UNMODIFIED reference classes, compiled by oracle/Makefile
// from the sources where they lie under /root/reference into oracle/_ref/libcsref.so:
//   src/api/fm_index.cpp, src/core/wavelet.cpp, src/core/bitvector.cpp  (+ their headers)
// It contains no algorithm of its own: every query entry point forwards to
// cs::FMIndex::count / cs::FMIndex::locate / cs::WaveletTree::rank / cs::BitVector::rank1.
//
// The only non-forwarding code is csref_inject*(): cs::build_sa_naive (src/core/sais.hpp:8-16)
// is O(n^2 log n), so for n beyond ~1e5 we fill FMIndex's private members ourselves, in the
// same order and with the same reference helpers as FMIndex::build_from_text
// (src/api/fm_index.cpp:16-69), from a suffix array supplied by the caller, and then call the
// reference's count()/locate() verbatim. This TU is compiled with -fno-access-control for that.
#include <cstring>  // the reference's serialization.hpp uses memset/memcpy without it
#include "api/fm_index.hpp"
#include "serialization/serialization.hpp"
#include "core/sais.hpp"
#include "core/bwt.hpp"

#include <array>
#include <atomic>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

using cs::FMIndex;

namespace {

// Same statements as src/api/fm_index.cpp:36-47 (C array) applied to an injected BWT.
void fill_C(FMIndex& idx) {
  idx.C_.assign(257, 0u);
  std::array<uint32_t, 256> freq{};
  freq.fill(0);
  for (unsigned char ch : idx.bwt_) freq[ch]++;
  uint32_t cum = 0;
  for (int c = 0; c < 256; ++c) { idx.C_[c] = cum; cum += freq[c]; }
  idx.C_[256] = cum;
}

void fill_wavelet(FMIndex& idx) {
  std::vector<uint8_t> bwt_bytes(idx.bwt_.begin(), idx.bwt_.end());
  idx.wavelet_.build(bwt_bytes);  // src/core/wavelet.cpp:14-53
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* csref_last_error() { return g_err.c_str(); }

// ---- FMIndex ---------------------------------------------------------------------------

// Verbatim cs::FMIndex::build_from_text (prints the reference's four [TIMER] lines to stderr).
void* csref_build_from_text(const uint8_t* text, uint64_t n, uint32_t ssa_stride) {
  cs::BuildParams p;
  p.ssa_stride = ssa_stride;
  std::string t(reinterpret_cast<const char*>(text), n);
  return new FMIndex(FMIndex::build_from_text(t, p));
}

// Full injection: text + SA -> every member build_from_text would have filled.
void* csref_inject(const uint8_t* text, uint64_t n, const uint32_t* sa, uint32_t ssa_stride) {
  auto* idx = new FMIndex();
  idx->text_.assign(reinterpret_cast<const char*>(text), n);
  idx->meta_.n = n;
  idx->sa_.assign(sa, sa + n);
  idx->bwt_ = cs::build_bwt_from_sa(idx->text_, idx->sa_);  // src/core/bwt.hpp:7-15
  fill_C(*idx);
  fill_wavelet(*idx);
  idx->ssa_.stride = ssa_stride;  // src/api/fm_index.cpp:57-65
  const size_t num_samples = (idx->sa_.size() + ssa_stride - 1) / ssa_stride;
  idx->ssa_.samples.resize(num_samples);
  for (size_t i = 0; i < idx->sa_.size(); ++i)
    if (i % ssa_stride == 0) idx->ssa_.samples[i / ssa_stride] = idx->sa_[i];
  return idx;
}

// Light injection for large n (no text_/sa_ copy): count() reads only meta_.n, C_, wavelet_;
// locate() additionally reads bwt_ and ssa_. samples may be null when only count() is used.
void* csref_inject_bwt(const uint8_t* bwt, uint64_t n, const uint32_t* samples, uint64_t nsamp,
                       uint32_t ssa_stride) {
  auto* idx = new FMIndex();
  idx->meta_.n = n;
  idx->bwt_.assign(reinterpret_cast<const char*>(bwt), n);
  fill_C(*idx);
  fill_wavelet(*idx);
  idx->ssa_.stride = ssa_stride;
  if (samples) idx->ssa_.samples.assign(samples, samples + nsamp);
  return idx;
}

// Fast injection for n ~ 1e9 (wavelet_.build alone would take minutes there): the eight packed
// bit planes are supplied by the caller and handed to the reference's own
// cs::BitVector::build_from_words (src/core/bitvector.cpp:98-159), which builds the reference's
// super/sub-block directory. tests/test_oracle_vs_reference.py checks that the resulting levels
// are identical to the ones cs::WaveletTree::build produces. words[l] must hold exactly
// ceil(n/64) u64 words. bwt/samples may be null when only count() will be called.
void* csref_inject_planes(uint64_t n, const uint64_t* const* words, const uint32_t* C257,
                          const uint8_t* bwt, const uint32_t* samples, uint64_t nsamp,
                          uint32_t ssa_stride) {
  auto* idx = new FMIndex();
  idx->meta_.n = n;
  idx->C_.assign(C257, C257 + 257);
  idx->wavelet_.n_ = n;
  const size_t nwords = (n + 63) / 64;
  if (n) {
    for (int l = 0; l < 8; ++l)
      idx->wavelet_.levels_[l].build_from_words(std::vector<uint64_t>(words[l], words[l] + nwords), n);
  }
  if (bwt) idx->bwt_.assign(reinterpret_cast<const char*>(bwt), n);
  idx->ssa_.stride = ssa_stride;
  if (samples) idx->ssa_.samples.assign(samples, samples + nsamp);
  return idx;
}

void csref_destroy(void* h) { delete static_cast<FMIndex*>(h); }

uint64_t csref_n(void* h) { return static_cast<FMIndex*>(h)->meta_.n; }

uint64_t csref_count(void* h, const uint8_t* pat, uint64_t m) {
  return static_cast<FMIndex*>(h)->count(std::string_view(reinterpret_cast<const char*>(pat), m));
}

// Returns the number of positions (written to out up to cap), or -1 if the reference threw
// (message available through csref_last_error()).
int64_t csref_locate(void* h, const uint8_t* pat, uint64_t m, uint64_t limit, uint64_t* out,
                     uint64_t cap) {
  try {
    auto v = static_cast<FMIndex*>(h)->locate(
        std::string_view(reinterpret_cast<const char*>(pat), m), static_cast<size_t>(limit));
    const uint64_t k = v.size() < cap ? v.size() : cap;
    if (k) std::memcpy(out, v.data(), k * sizeof(uint64_t));
    return static_cast<int64_t>(v.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// count() over a packed batch (bytes + offs[npat+1]) sliced across nthreads std::threads.
// The methods are const with no mutable state, so concurrent readers are safe.
void csref_count_many(void* h, const uint8_t* bytes, const uint64_t* offs, uint64_t npat,
                      uint64_t* out, int nthreads) {
  auto* idx = static_cast<FMIndex*>(h);
  if (nthreads < 1) nthreads = 1;
  std::atomic<uint64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const uint64_t q = next.fetch_add(1);
      if (q >= npat) break;
      out[q] = idx->count(std::string_view(reinterpret_cast<const char*>(bytes + offs[q]),
                                           offs[q + 1] - offs[q]));
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

// locate() over a packed batch; out_n[q] = number of positions or -1 if the reference threw;
// returns the total number of positions (positions themselves are discarded unless out_pos
// is non-null, in which case query q's positions are written at out_pos + q*limit).
uint64_t csref_locate_many(void* h, const uint8_t* bytes, const uint64_t* offs, uint64_t npat,
                           uint64_t limit, int64_t* out_n, uint64_t* out_pos, int nthreads) {
  auto* idx = static_cast<FMIndex*>(h);
  if (nthreads < 1) nthreads = 1;
  std::atomic<uint64_t> next{0}, total{0};
  auto work = [&]() {
    for (;;) {
      const uint64_t q = next.fetch_add(1);
      if (q >= npat) break;
      try {
        auto v = idx->locate(std::string_view(reinterpret_cast<const char*>(bytes + offs[q]),
                                              offs[q + 1] - offs[q]),
                             static_cast<size_t>(limit));
        out_n[q] = static_cast<int64_t>(v.size());
        total.fetch_add(v.size());
        if (out_pos && !v.empty())
          std::memcpy(out_pos + q * limit, v.data(), v.size() * sizeof(uint64_t));
      } catch (const std::exception&) {
        out_n[q] = -1;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return total.load();
}

uint64_t csref_extract(void* h, uint64_t pos, uint64_t len, uint8_t* out) {
  std::string s = static_cast<FMIndex*>(h)->extract(pos, len);
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return s.size();
}

// Dumps of the private build products (golden-vector generation / oracle pinning).
void csref_get_sa(void* h, uint32_t* out) {
  auto* idx = static_cast<FMIndex*>(h);
  if (!idx->sa_.empty()) std::memcpy(out, idx->sa_.data(), idx->sa_.size() * 4);
}
void csref_get_bwt(void* h, uint8_t* out) {
  auto* idx = static_cast<FMIndex*>(h);
  if (!idx->bwt_.empty()) std::memcpy(out, idx->bwt_.data(), idx->bwt_.size());
}
void csref_get_C(void* h, uint32_t* out257) {
  auto* idx = static_cast<FMIndex*>(h);
  std::memcpy(out257, idx->C_.data(), 257 * 4);
}
uint64_t csref_ssa_size(void* h) { return static_cast<FMIndex*>(h)->ssa_.samples.size(); }
void csref_get_ssa(void* h, uint32_t* out) {
  auto* idx = static_cast<FMIndex*>(h);
  if (!idx->ssa_.samples.empty())
    std::memcpy(out, idx->ssa_.samples.data(), idx->ssa_.samples.size() * 4);
}
uint64_t csref_occ(void* h, uint8_t c, uint64_t i) { return static_cast<FMIndex*>(h)->occ(c, i); }
uint64_t csref_LF(void* h, uint64_t i) { return static_cast<FMIndex*>(h)->LF(i); }

// ---- standalone cs::build_sa_naive -------------------------------------------------------
void csref_build_sa_naive(const uint8_t* text, uint64_t n, uint32_t* out) {
  std::string t(reinterpret_cast<const char*>(text), n);
  auto sa = cs::build_sa_naive(t);
  if (n) std::memcpy(out, sa.data(), n * 4);
}

// ---- standalone cs::WaveletTree ------------------------------------------------------------
void* csref_wt_build(const uint8_t* seq, uint64_t n) {
  auto* wt = new cs::WaveletTree();
  wt->build(std::vector<uint8_t>(seq, seq + n));
  return wt;
}
void csref_wt_destroy(void* w) { delete static_cast<cs::WaveletTree*>(w); }
uint64_t csref_wt_rank(void* w, uint8_t c, uint64_t i) {
  return static_cast<cs::WaveletTree*>(w)->rank(c, i);
}
uint8_t csref_wt_access(void* w, uint64_t i) { return static_cast<cs::WaveletTree*>(w)->access(i); }
// level directory dumps (BitVector::bits/super_blocks/sub_blocks, bitvector.hpp:88-92)
uint64_t csref_wt_level_sizes(void* w, int level, uint64_t* nwords, uint64_t* nsuper,
                              uint64_t* nsub) {
  const cs::BitVector& bv = static_cast<cs::WaveletTree*>(w)->levels_[level];
  *nwords = bv.bits().size();
  *nsuper = bv.super_blocks().size();
  *nsub = bv.sub_blocks().size();
  return bv.size();
}
void csref_wt_level_dump(void* w, int level, uint64_t* words, uint32_t* supers, uint16_t* subs) {
  const cs::BitVector& bv = static_cast<cs::WaveletTree*>(w)->levels_[level];
  if (!bv.bits().empty()) std::memcpy(words, bv.bits().data(), bv.bits().size() * 8);
  if (!bv.super_blocks().empty())
    std::memcpy(supers, bv.super_blocks().data(), bv.super_blocks().size() * 4);
  if (!bv.sub_blocks().empty())
    std::memcpy(subs, bv.sub_blocks().data(), bv.sub_blocks().size() * 2);
}

// ---- standalone cs::BitVector ----------------------------------------------------------------
void* csref_bv_build(const uint8_t* bits01, uint64_t n) {
  auto* bv = new cs::BitVector();
  bv->build(std::vector<uint8_t>(bits01, bits01 + n));
  return bv;
}
void* csref_bv_build_from_words(const uint64_t* words, uint64_t nwords, uint64_t nbits) {
  auto* bv = new cs::BitVector();
  bv->build_from_words(std::vector<uint64_t>(words, words + nwords), nbits);
  return bv;
}
void csref_bv_destroy(void* b) { delete static_cast<cs::BitVector*>(b); }
uint64_t csref_bv_rank1(void* b, uint64_t i) { return static_cast<cs::BitVector*>(b)->rank1(i); }
uint64_t csref_bv_rank0(void* b, uint64_t i) { return static_cast<cs::BitVector*>(b)->rank0(i); }
uint8_t csref_bv_get(void* b, uint64_t i) { return static_cast<cs::BitVector*>(b)->get(i); }
uint64_t csref_bv_size(void* b) { return static_cast<cs::BitVector*>(b)->size(); }

// ---- cs::IndexReader / cs::IndexWriter (src/serialization/serialization.cpp) ----------------------
// The reader is the format oracle for .csidx files. The writer is only usable when no padding is
// ever needed (its align_to never terminates otherwise, serialization.cpp:44-54).
void* csref_reader_open(const char* path) {
  try {
    return new cs::IndexReader(path);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void csref_reader_close(void* r) { delete static_cast<cs::IndexReader*>(r); }
int csref_reader_header(void* r, uint32_t* flags, uint64_t* text_len, uint64_t* offsets8) {
  const cs::IndexHeader* h = static_cast<cs::IndexReader*>(r)->header();
  if (!h || !h->is_valid()) return 1;
  *flags = h->flags;
  *text_len = h->text_len;
  std::memcpy(offsets8, h->offsets, 64);
  return 0;
}
const void* csref_reader_section(void* r, int which, uint64_t* count, uint32_t* stride) {
  auto* rd = static_cast<cs::IndexReader*>(r);
  size_t n = 0;
  uint32_t st = 0;
  const void* p = nullptr;
  switch (which) {
    case 1: p = rd->get_text(&n); break;
    case 2: p = rd->get_bwt(&n); break;
    case 3: p = rd->get_c_array(&n); break;
    case 4: p = rd->get_ssa(&n, &st); break;
    case 5: p = rd->get_wavelet(&n); break;
    case 6: p = rd->get_veb_layout(&n); break;
    default: break;
  }
  *count = n;
  if (stride) *stride = st;
  return p;
}
// header + TEXT + BWT + C_ARRAY with the reference writer (caller guarantees no padding is needed)
int csref_writer_simple(const char* path, uint32_t flags, const uint8_t* text, uint64_t ntext, const uint8_t* bwt,
                        uint64_t nbwt, const uint32_t* c, uint64_t nc) {
  try {
    cs::IndexWriter w(path);
    w.write_header(flags, ntext);
    w.write_text(std::string(reinterpret_cast<const char*>(text), ntext));
    w.write_bwt(std::vector<uint8_t>(bwt, bwt + nbwt));
    w.write_c_array(std::vector<uint32_t>(c, c + nc));
    w.finalize();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

int csref_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
