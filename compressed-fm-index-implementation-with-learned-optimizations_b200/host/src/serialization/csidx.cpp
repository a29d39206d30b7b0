// host/src/serialization/csidx.cpp — see csidx.hpp.
#include "csidx.hpp"

#include <cstring>
#include <fstream>
#include <stdexcept>

namespace cs {

namespace {

constexpr uint64_t kFooter = 0x444E4553435300ULL;  // serialization.cpp:139
enum { SEC_HEADER = 0, SEC_TEXT, SEC_BWT, SEC_C, SEC_SSA, SEC_WAVELET, SEC_LAYOUT, SEC_FOOTER, NSEC };

#pragma pack(push, 1)
struct Header {  // serialization.hpp:64-81
  char magic[8];
  uint16_t version;
  uint16_t reserved1;
  uint32_t flags;
  uint64_t text_len;
  uint64_t offsets[NSEC];
};
#pragma pack(pop)
static_assert(sizeof(Header) == 88, "IndexHeader is 88 bytes (serialization.hpp:83)");

class Out {
public:
  explicit Out(const std::string& path) : f_(path, std::ios::binary | std::ios::trunc) {
    if (!f_) throw std::runtime_error("Failed to open file for writing: " + path);
    const char zeros[sizeof(Header)] = {0};
    raw(zeros, sizeof zeros);  // header is written last (serialization.cpp:143-144)
  }
  void raw(const void* p, size_t n) {
    if (n) f_.write(static_cast<const char*>(p), static_cast<std::streamsize>(n));
    if (!f_) throw std::runtime_error("Write failed");
    off_ += n;
  }
  void align(size_t a) {  // what IndexWriter::align_to means to do: zero padding
    static const char zeros[4096] = {0};
    const size_t pad = (a - off_ % a) % a;
    raw(zeros, pad);
  }
  template <class T> void array(const std::vector<T>& v) {  // serialization.hpp:113-120
    const uint64_t count = v.size();
    raw(&count, 8);
    raw(v.data(), v.size() * sizeof(T));
  }
  void finish(const Header& h) {
    f_.seekp(0);
    f_.write(reinterpret_cast<const char*>(&h), sizeof h);
    f_.close();
    if (!f_) throw std::runtime_error("Write failed");
  }
  uint64_t off() const { return off_; }

private:
  std::ofstream f_;
  uint64_t off_ = 0;
};

}  // namespace

void write_csidx(const std::string& path, const CsidxSections& s) {
  Out out(path);
  Header h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "CSIDX", 5);
  h.version = 1;
  h.flags = s.flags | (s.device_blob.empty() ? 0u : CSIDX_FLAG_DEVICE_BLOB);
  h.text_len = s.text_len ? s.text_len : (s.has_text ? s.text.size() : s.bwt.size());
  if (s.has_text) {  // serialization.cpp:70-77
    out.align(8);
    h.offsets[SEC_TEXT] = out.off();
    const uint64_t len = s.text.size();
    out.raw(&len, 8);
    out.raw(s.text.data(), s.text.size());
  }
  if (!s.bwt.empty()) {  // :79-83
    out.align(8);
    h.offsets[SEC_BWT] = out.off();
    out.array(s.bwt);
  }
  if (!s.c_array.empty()) {  // :85-89
    out.align(8);
    h.offsets[SEC_C] = out.off();
    out.array(s.c_array);
  }
  if (s.has_ssa) {  // :91-101 (the reader expects the array at offset + 8, :313)
    out.align(8);
    h.offsets[SEC_SSA] = out.off();
    out.raw(&s.ssa_stride, 4);
    out.align(8);
    out.array(s.ssa);
  }
  {  // :103-118 — present but empty
    out.align(8);
    h.offsets[SEC_WAVELET] = out.off();
    const uint64_t levels = 0;
    out.raw(&levels, 8);
    out.array(std::vector<uint64_t>());
    out.array(std::vector<uint32_t>());
    out.array(std::vector<uint16_t>());
  }
  if (!s.device_blob.empty()) {  // :120-132
    out.align(4096);
    h.offsets[SEC_LAYOUT] = out.off();
    out.array(s.device_blob);
  }
  out.align(8);  // :135-139
  h.offsets[SEC_FOOTER] = out.off();
  out.raw(&kFooter, 8);
  out.finish(h);
}

CsidxSections read_csidx(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("Failed to open file: " + path);
  const uint64_t size = static_cast<uint64_t>(f.tellg());
  if (size < sizeof(Header) + 8) throw std::runtime_error("csidx: file too small");
  f.seekg(0);
  Header h;
  f.read(reinterpret_cast<char*>(&h), sizeof h);
  if (std::memcmp(h.magic, "CSIDX", 5) != 0 || h.version != 1) throw std::runtime_error("csidx: bad magic or version");
  auto at = [&](uint64_t off, void* p, uint64_t n) {
    if (off > size || n > size - off) throw std::runtime_error("csidx: section out of bounds");
    f.seekg(static_cast<std::streamoff>(off));
    f.read(static_cast<char*>(p), static_cast<std::streamsize>(n));
    if (!f) throw std::runtime_error("csidx: read failed");
  };
  uint64_t footer = 0;
  at(h.offsets[SEC_FOOTER], &footer, 8);
  if (h.offsets[SEC_FOOTER] == 0 || footer != kFooter) throw std::runtime_error("csidx: missing footer");
  auto count_at = [&](uint64_t off) { uint64_t c = 0; at(off, &c, 8); return c; };
  CsidxSections s;
  s.flags = h.flags;
  s.text_len = h.text_len;
  if (uint64_t o = h.offsets[SEC_TEXT]) {
    const uint64_t n = count_at(o);
    if (n > size) throw std::runtime_error("csidx: section out of bounds");
    s.has_text = true;
    s.text.resize(n);
    at(o + 8, s.text.data(), n);
  }
  if (uint64_t o = h.offsets[SEC_BWT]) {
    const uint64_t n = count_at(o);
    if (n > size) throw std::runtime_error("csidx: section out of bounds");
    s.bwt.resize(n);
    at(o + 8, s.bwt.data(), n);
  }
  if (uint64_t o = h.offsets[SEC_C]) {
    const uint64_t n = count_at(o);
    if (n > size / 4) throw std::runtime_error("csidx: section out of bounds");
    s.c_array.resize(n);
    at(o + 8, s.c_array.data(), n * 4);
  }
  if (uint64_t o = h.offsets[SEC_SSA]) {
    at(o, &s.ssa_stride, 4);
    const uint64_t n = count_at(o + 8);
    if (n > size / 4) throw std::runtime_error("csidx: section out of bounds");
    s.has_ssa = true;
    s.ssa.resize(n);
    at(o + 16, s.ssa.data(), n * 4);
  }
  if (uint64_t o = h.offsets[SEC_LAYOUT]) {
    if (h.flags & CSIDX_FLAG_DEVICE_BLOB) {
      const uint64_t n = count_at(o);
      if (n > size) throw std::runtime_error("csidx: section out of bounds");
      s.device_blob.resize(n);
      at(o + 8, s.device_blob.data(), n);
    }
  }
  return s;
}

}  // namespace cs

// C entry points for the format tests (ctypes): write from caller arrays, read into caller arrays.
extern "C" {

__attribute__((visibility("default"))) int cs_b200_csidx_write(const char* path, uint32_t flags, const uint8_t* text,
                                                                uint64_t ntext, int has_text, const uint8_t* bwt,
                                                                uint64_t nbwt, const uint32_t* c, uint64_t nc,
                                                                const uint32_t* ssa, uint64_t nssa, uint32_t stride,
                                                                int has_ssa, const uint8_t* blob, uint64_t nblob) {
  try {
    cs::CsidxSections s;
    s.flags = flags;
    s.has_text = has_text != 0;
    if (has_text) s.text.assign(reinterpret_cast<const char*>(text), ntext);
    s.bwt.assign(bwt, bwt + nbwt);
    s.c_array.assign(c, c + nc);
    s.has_ssa = has_ssa != 0;
    s.ssa.assign(ssa, ssa + nssa);
    s.ssa_stride = stride;
    s.device_blob.assign(blob, blob + nblob);
    s.text_len = has_text ? ntext : nbwt;
    cs::write_csidx(path, s);
    return 0;
  } catch (...) {
    return 1;
  }
}

// Sizes first (out8 = {ntext, nbwt, nc, nssa, stride, nblob, flags, text_len}), then the payloads
// into caller buffers (any of which may be null).
__attribute__((visibility("default"))) int cs_b200_csidx_read(const char* path, uint64_t* out8, uint8_t* text, uint8_t* bwt,
                                                               uint32_t* c, uint32_t* ssa, uint8_t* blob) {
  try {
    const cs::CsidxSections s = cs::read_csidx(path);
    out8[0] = s.text.size(); out8[1] = s.bwt.size(); out8[2] = s.c_array.size(); out8[3] = s.ssa.size();
    out8[4] = s.ssa_stride; out8[5] = s.device_blob.size(); out8[6] = s.flags; out8[7] = s.text_len;
    if (text && !s.text.empty()) std::memcpy(text, s.text.data(), s.text.size());
    if (bwt && !s.bwt.empty()) std::memcpy(bwt, s.bwt.data(), s.bwt.size());
    if (c && !s.c_array.empty()) std::memcpy(c, s.c_array.data(), s.c_array.size() * 4);
    if (ssa && !s.ssa.empty()) std::memcpy(ssa, s.ssa.data(), s.ssa.size() * 4);
    if (blob && !s.device_blob.empty()) std::memcpy(blob, s.device_blob.data(), s.device_blob.size());
    return 0;
  } catch (...) {
    return 1;
  }
}
}
