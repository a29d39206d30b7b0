// csfm_api.cu — the extern "C" surface declared in include/csfm.h.
//
// Host logic only: argument checks, host<->device staging for the host-pointer entry points,
// workspace reuse, accounting. No algorithm lives here and nothing here can run a query on the
// CPU: every path ends in a kernel launch from csfm_query.cu / csfm_build.cu / csfm_sa.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "csfm_dna.cuh"
#include "csfm_host.hpp"
#include "csfm_kernels.cuh"

namespace csfm {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
PhaseTimer::PhaseTimer(const char* w) : on(std::getenv("CSFM_BUILD_TIMERS") != nullptr), t0(0), what(w) {
  if (on) {
    cudaDeviceSynchronize();
    t0 = now_s();
  }
}
void PhaseTimer::mark(const char* phase) {
  if (!on) return;
  cudaDeviceSynchronize();
  const double t = now_s();
  std::fprintf(stderr, "[csfm build] %s / %s: %.1f ms\n", what, phase, 1e3 * (t - t0));
  t0 = t;
}

int DeviceBuffer::ensure(size_t bytes) {
  if (bytes <= cap) return CSFM_OK;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) return fail(CSFM_ERR_NOMEM, std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e));
  cap = want;
  return CSFM_OK;
}
void DeviceBuffer::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

DeviceGuard::DeviceGuard(int dev) {
  if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
  if (prev == dev) {  // the common case: nothing to switch, nothing to restore (two driver calls less per query)
    prev = -1;
    ok = true;
    return;
  }
  ok = (cudaSetDevice(dev) == cudaSuccess);
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0) cudaSetDevice(prev);
}

static int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(CSFM_ERR_CUDA, std::string("no usable CUDA device (this engine has no CPU fallback): ") +
                                   cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(CSFM_ERR_INVALID, "device ordinal out of range");
  return CSFM_OK;
}

static uint32_t stride_of(const csfm_params* p) { return p ? p->ssa_stride : 32u; }

// Host-pointer entry points check what the kernels trust: offsets must not decrease (a decreasing pair would
// give a pattern length near 2^64 and a walk far outside the staged bytes). One pass over npat + 1 words.
template <class T>
static bool offsets_monotone(const T* offs, uint64_t npat) {
  T bad = 0;
  for (uint64_t i = 0; i < npat; ++i) bad |= (T)(offs[i + 1] < offs[i]);
  return bad == 0;
}

}  // namespace csfm

using namespace csfm;

extern "C" {

const char* csfm_last_error(void) { return g_last_error.c_str(); }
const char* csfm_version(void) { return "csfm-b200 0.2 (sm_100a)"; }

int csfm_device_count(int* count) {
  if (!count) return fail(CSFM_ERR_INVALID, "count is null");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(CSFM_ERR_CUDA, cudaGetErrorString(e));
  }
  return CSFM_OK;
}

int csfm_build_from_text_device(const uint8_t* d_text, uint64_t n, const csfm_params* params, int device,
                                uint32_t flags, csfm_index** out) {
  if (!out) return fail(CSFM_ERR_INVALID, "out is null");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  if (n && !d_text) return fail(CSFM_ERR_INVALID, "text is null");
  if (n > kMaxN) return fail(CSFM_ERR_TOO_LARGE, "text length must be < 2^32 - 1");
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  const uint32_t stride = stride_of(params);
  uint8_t* d_bwt = nullptr;
  uint32_t *d_ssa = nullptr, *d_sa = nullptr;
  uint64_t nsamp = 0;
  const bool want_sa = (flags & CSFM_BUILD_KEEP_SA) || !((flags & CSFM_BUILD_NO_TEXT_CHECK) || (flags & CSFM_BUILD_LAYOUT_BINARY64));
  uint32_t sa_rounds = 0, sa_passes = 0;
  uint64_t sa_pair_passes = 0;
  rc = build_sa_bwt_device(d_text, n, stride, nullptr, &d_bwt, &d_ssa, &nsamp, want_sa ? &d_sa : nullptr, &sa_rounds, &sa_passes,
                           &sa_pair_passes);
  if (rc) return rc;
  rc = index_from_device_bwt(d_bwt, n, d_ssa, nsamp, stride, device, flags, out, d_text, d_sa);
  cudaFree(d_bwt);
  cudaFree(d_ssa);
  if (rc) {
    cudaFree(d_sa);
    return rc;
  }
  (*out)->sa_rounds = sa_rounds;
  (*out)->sa_radix_passes = sa_passes;
  (*out)->sa_pair_passes = sa_pair_passes;
  if (flags & CSFM_BUILD_KEEP_SA)
    (*out)->d_sa = d_sa;
  else
    cudaFree(d_sa);
  return CSFM_OK;
}

int csfm_build_from_text(const uint8_t* text, uint64_t n, const csfm_params* params, int device, uint32_t flags,
                         csfm_index** out) {
  if (!out) return fail(CSFM_ERR_INVALID, "out is null");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  if (n && !text) return fail(CSFM_ERR_INVALID, "text is null");
  if (n > kMaxN) return fail(CSFM_ERR_TOO_LARGE, "text length must be < 2^32 - 1");
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  uint8_t* d_text = nullptr;
  CSFM_CUDA(cudaMalloc(&d_text, n ? n : 1));
  cudaError_t e = cudaMemcpy(d_text, text, n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_text);
    return fail(CSFM_ERR_CUDA, cudaGetErrorString(e));
  }
  rc = csfm_build_from_text_device(d_text, n, params, device, flags, out);
  cudaFree(d_text);
  return rc;
}

int csfm_build_from_parts(const uint8_t* bwt, uint64_t n, const uint32_t* ssa, uint64_t nsamp, uint32_t ssa_stride,
                          int device, uint32_t flags, csfm_index** out) {
  if (!out) return fail(CSFM_ERR_INVALID, "out is null");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  if ((n && !bwt) || (nsamp && !ssa)) return fail(CSFM_ERR_INVALID, "null input");
  if (ssa_stride == 0) return fail(CSFM_ERR_INVALID, "ssa_stride must be > 0");
  if (nsamp != (n + ssa_stride - 1) / ssa_stride)
    return fail(CSFM_ERR_INVALID, "nsamp must equal ceil(n / ssa_stride)");
  if (n > kMaxN) return fail(CSFM_ERR_TOO_LARGE, "text length must be < 2^32 - 1");
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  uint8_t* d_bwt = nullptr;
  uint32_t* d_ssa = nullptr;
  CSFM_CUDA(cudaMalloc(&d_bwt, n ? n : 1));
  cudaError_t e = cudaMalloc(&d_ssa, nsamp ? nsamp * 4 : 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_bwt, bwt, n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nsamp) e = cudaMemcpy(d_ssa, ssa, nsamp * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_bwt);
    cudaFree(d_ssa);
    return fail(CSFM_ERR_CUDA, cudaGetErrorString(e));
  }
  rc = index_from_device_bwt(d_bwt, n, d_ssa, nsamp, ssa_stride, device, flags, out);
  cudaFree(d_bwt);
  cudaFree(d_ssa);
  return rc;
}

void csfm_destroy(csfm_index* idx) {
  if (!idx) return;
  DeviceGuard g(idx->device);
  if (idx->stream) cudaStreamSynchronize(idx->stream);
  for (auto& s : idx->aux_stream)
    if (s) cudaStreamSynchronize(s);
  for (auto& sl : idx->async_slot) {
    if (sl.stream) {
      cudaStreamSynchronize(sl.stream);
      cudaStreamDestroy(sl.stream);
    }
    sl.in.release();
    sl.out.release();
    sl.scan.release();
  }
  for (auto& qs : idx->qslot) {
    qs.buf.release();
    if (qs.done) cudaEventDestroy(qs.done);
  }
  if (idx->owns_blob && idx->d_blob) cudaFree(idx->d_blob);
  if (idx->d_sa) cudaFree(idx->d_sa);
  if (idx->d_text_cache) cudaFree(idx->d_text_cache);
  idx->ws_in.release();
  idx->ws_out.release();
  idx->ws_tmp.release();
  idx->ws_scan.release();
  idx->ws_pos.release();
  if (idx->d_counters) cudaFree(idx->d_counters);
  if (idx->h_pinned) cudaFreeHost(idx->h_pinned);
  if (idx->ev0) cudaEventDestroy(idx->ev0);
  if (idx->ev1) cudaEventDestroy(idx->ev1);
  if (idx->ev_ws_tmp) cudaEventDestroy(idx->ev_ws_tmp);
  for (auto& e : idx->ev_slice)
    if (e) cudaEventDestroy(e);
  if (idx->stream) cudaStreamDestroy(idx->stream);
  for (auto& s : idx->aux_stream)
    if (s) cudaStreamDestroy(s);
  delete idx;
}

int csfm_info(const csfm_index* idx, csfm_index_info* out) {
  if (!idx || !out) return fail(CSFM_ERR_INVALID, "null argument");
  std::memset(out, 0, sizeof *out);
  out->n = idx->h.n;
  out->sigma = idx->h.sigma;
  out->levels = idx->h.levels;
  out->ssa_stride = idx->h.stride;
  out->device = (uint32_t)idx->device;
  out->nsamp = idx->h.nsamp;
  out->blocks_per_level = idx->h.nblk;
  out->blob_bytes = idx->blob_bytes;
  out->has_sa = idx->d_sa != nullptr;
  out->layout = idx->h.layout;
  out->line_bytes = idx->h.layout == kLayoutNibble128 ? kLine2Bytes : kLineBytes;  // layouts 1 and 3: 64-byte lines
  out->kmer_k = idx->view.kmer_k;
  out->text_check = idx->view.text != nullptr;
  out->half_table = idx->view.kmer_hi != nullptr;
  out->sa_rounds = idx->sa_rounds;
  out->sa_radix_passes = idx->sa_radix_passes;
  out->sa_pair_passes = idx->sa_pair_passes;
  out->position_samples = idx->h.layout == kLayoutDna64 && idx->h.marked;
  return CSFM_OK;
}

int csfm_get_C(const csfm_index* idx, uint32_t C[257]) {
  if (!idx || !C) return fail(CSFM_ERR_INVALID, "null argument");
  std::memcpy(C, idx->h.C, 257 * 4);
  return CSFM_OK;
}

int csfm_get_ssa(const csfm_index* idx, uint32_t* out) {
  if (!idx || (!out && idx->h.nsamp)) return fail(CSFM_ERR_INVALID, "null argument");
  DeviceGuard g(idx->device);
  if (idx->h.nsamp)
    CSFM_CUDA(cudaMemcpy(out, idx->d_blob + idx->h.off_ssa, idx->h.nsamp * 4, cudaMemcpyDeviceToHost));
  return CSFM_OK;
}

int csfm_get_sa(const csfm_index* idx, uint32_t* out) {
  if (!idx || (!out && idx->h.n)) return fail(CSFM_ERR_INVALID, "null argument");
  if (idx->h.n && !idx->d_sa) return fail(CSFM_ERR_INVALID, "suffix array not resident (build with CSFM_BUILD_KEEP_SA)");
  DeviceGuard g(idx->device);
  if (idx->h.n) CSFM_CUDA(cudaMemcpy(out, idx->d_sa, idx->h.n * 4, cudaMemcpyDeviceToHost));
  return CSFM_OK;
}

int csfm_sa_device(const csfm_index* idx, const uint32_t** d_sa) {
  if (!idx || !d_sa) return fail(CSFM_ERR_INVALID, "null argument");
  if (idx->h.n && !idx->d_sa) return fail(CSFM_ERR_INVALID, "suffix array not resident (build with CSFM_BUILD_KEEP_SA)");
  *d_sa = idx->d_sa;
  return CSFM_OK;
}

int csfm_release_sa(csfm_index* idx) {
  if (!idx) return fail(CSFM_ERR_INVALID, "null argument");
  DeviceGuard g(idx->device);
  if (idx->d_sa) cudaFree(idx->d_sa);
  idx->d_sa = nullptr;
  return CSFM_OK;
}

int csfm_extract_bwt(const csfm_index* cidx, uint8_t* out) {
  csfm_index* idx = const_cast<csfm_index*>(cidx);
  if (!idx || (!out && idx->h.n)) return fail(CSFM_ERR_INVALID, "null argument");
  if (idx->h.n == 0) return CSFM_OK;
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  int rc = idx->ws_out.ensure(idx->h.n);
  if (rc) return rc;
  rc = extract_bwt_device(idx, idx->ws_out.as<uint8_t>(), idx->stream);
  if (rc) return rc;
  CSFM_CUDA(cudaMemcpyAsync(out, idx->ws_out.p, idx->h.n, cudaMemcpyDeviceToHost, idx->stream));
  CSFM_CUDA(cudaStreamSynchronize(idx->stream));
  return CSFM_OK;
}

// cs::FMIndex::extract (fm_index.cpp:163-167) without a host copy of the text: from the blob's text section when
// it has one, else from a device copy of the text that is rebuilt ONCE out of the index itself (every sampled row
// walks LF to the next sampled row and writes the BWT symbols it passes: n LF steps in all).
int csfm_extract(csfm_index* idx, uint64_t pos, uint64_t len, uint8_t* out, uint64_t* got) {
  if (!idx || !got || (len && !out)) return fail(CSFM_ERR_INVALID, "null argument");
  *got = 0;
  const uint64_t n = idx->h.n;
  if (pos >= n || len == 0) return CSFM_OK;  // fm_index.cpp:164
  len = std::min(len, n - pos);
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  const uint8_t* src = idx->h.off_text ? idx->d_blob + idx->h.off_text : idx->d_text_cache;
  if (!src) {
    if (idx->h.layout == kLayoutBinary64) return fail(CSFM_ERR_INVALID, "extract from the index needs layout 2 or 3 (or a text section)");
    // LF(row of suffix p) is the row of suffix p - 1 only when the text ends in a unique smallest byte (else the
    // reference's BWT is not a rotation BWT: its locate() over-counts or throws there too): the smallest byte present
    // occurs once, and row 0 — always sampled — is the suffix that starts at n - 1
    {
      int c = 0;
      while (c < 256 && idx->h.C[c + 1] == idx->h.C[c]) ++c;
      uint32_t sa0 = 0;
      CSFM_CUDA(cudaMemcpy(&sa0, idx->d_blob + idx->h.off_ssa, 4, cudaMemcpyDeviceToHost));
      if (c == 256 || idx->h.C[c + 1] - idx->h.C[c] != 1 || (uint64_t)sa0 != n - 1)
        return fail(CSFM_ERR_INVALID, "extract: the text cannot be rebuilt from this index (it does not end in a unique smallest byte)");
    }
    uint8_t* d = nullptr;
    CSFM_CUDA(cudaMalloc(&d, n));
    unsigned long long* ctr = next_counter_slot(idx);
    cudaError_t e = cudaMemsetAsync(ctr, 0, kCounterWords * sizeof(unsigned long long), idx->stream);
    if (e == cudaSuccess) {
      if (idx->h.layout == kLayoutDna64) launch_untext3(idx->view, d, ctr, ctr + 1, idx->num_sms, idx->stream);
      else launch_untext2(idx->view, d, ctr, ctr + 1, idx->num_sms, idx->stream);
      e = cudaGetLastError();
    }
    unsigned long long written = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&written, ctr + 1, 8, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess || written != n) {
      cudaFree(d);
      if (e != cudaSuccess) return fail(CSFM_ERR_CUDA, std::string("extract: ") + cudaGetErrorString(e));
      return fail(CSFM_ERR_INVALID, "extract: the text cannot be rebuilt from this index (LF cycles without a sampled row)");
    }
    idx->d_text_cache = d;
    src = d;
  }
  CSFM_CUDA(cudaMemcpyAsync(out, src + pos, len, cudaMemcpyDeviceToHost, idx->stream));
  CSFM_CUDA(cudaStreamSynchronize(idx->stream));
  *got = len;
  return CSFM_OK;
}

int csfm_blob(const csfm_index* idx, const void** d_blob, uint64_t* bytes) {
  if (!idx || !d_blob || !bytes) return fail(CSFM_ERR_INVALID, "null argument");
  *d_blob = idx->d_blob;
  *bytes = idx->blob_bytes;
  return CSFM_OK;
}

int csfm_attach_blob(void* d_blob, uint64_t bytes, int device, int take_ownership, csfm_index** out) {
  if (!out || !d_blob) return fail(CSFM_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  if (bytes < kHeaderBytes) return fail(CSFM_ERR_FORMAT, "blob smaller than its header");
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  BlobHeader h;
  CSFM_CUDA(cudaMemcpy(&h, d_blob, sizeof h, cudaMemcpyDeviceToHost));
  if (std::memcmp(h.magic, "CSFMDEV1", 8) != 0 || h.version != 3) return fail(CSFM_ERR_FORMAT, "bad blob magic/version");
  const bool nib = h.layout == kLayoutNibble128;
  const bool dna = h.layout == kLayoutDna64;
  const uint64_t per_line = dna ? (h.marked ? kSymsPerLine3M : kSymsPerLine3) : nib ? kSymsPerLine : kPayloadBits, line_bytes = nib ? kLine2Bytes : kLineBytes;
  if (h.marked > 1 || (h.marked && !dna) || (h.marked != 0) != (h.off_psamp != 0) ||
      (h.marked && (h.off_psamp % 4 || h.off_psamp < h.off_ssa + h.nsamp * 4 || h.off_psamp > h.total_bytes || h.nsamp * 4 > h.total_bytes - h.off_psamp)))
    return fail(CSFM_ERR_FORMAT, "inconsistent position samples in blob header");
  if ((h.layout != kLayoutNibble128 && h.layout != kLayoutBinary64 && h.layout != kLayoutDna64) || h.total_bytes > bytes || h.levels == 0 ||
      h.levels > (nib ? 2u : dna ? 1u : kMaxLevels) || h.nblk != h.n / per_line + 1 || h.off_levels < sizeof(BlobHeader) ||
      h.off_ssa + h.nsamp * 4 > h.total_bytes || h.stride == 0 ||
      h.off_levels + (uint64_t)h.levels * h.level_stride > h.off_ssa || h.level_stride < h.nblk * line_bytes ||
      h.nsamp != (h.n + h.stride - 1) / h.stride)
    return fail(CSFM_ERR_FORMAT, "inconsistent blob header");
  if (h.off_text) {
    if (!nib || h.n < 2 || h.dense_shift > 16 || h.off_text % 16 || h.off_text + h.n + 64 > h.off_dense ||
        h.off_dense + (((h.n - 1) >> h.dense_shift) + 1) * 4 > h.total_bytes || h.off_text < h.off_ssa + h.nsamp * 4)
      return fail(CSFM_ERR_FORMAT, "inconsistent text sections in blob header");
  }
  if (h.kmer_k) {
    uint64_t entries = 1;
    for (uint32_t i = 0; i < h.kmer_k && entries <= (1ull << 40); ++i) entries *= h.kmer_radix;
    const uint64_t table_bytes = h.kmer_tiled ? (entries + 1) * 4 : entries * 8;
    if (!(nib || dna) || (dna && (h.kmer_tiled || h.off_kmer_hi)) || h.kmer_k > 16 || h.kmer_radix < 2 || h.kmer_radix > 256 || h.off_kmer < h.off_ssa + h.nsamp * 4 ||
        (h.marked && h.off_kmer < h.off_psamp + h.nsamp * 4) ||
        h.kmer_tiled > 1 || (h.kmer_tiled && !h.off_text) || h.off_kmer + table_bytes > h.total_bytes)
      return fail(CSFM_ERR_FORMAT, "inconsistent k-mer table in blob header");
    if (h.off_kmer_hi && (h.levels != 2 || h.off_kmer_hi < h.off_kmer + table_bytes || h.off_kmer_hi + entries * 128 > h.total_bytes))
      return fail(CSFM_ERR_FORMAT, "inconsistent half-step table in blob header");
  }
  // The kernels turn header tables into line addresses without further checks: a corrupt or crafted
  // .csidx must fail here, not as an illegal address inside a kernel (which poisons the context).
  {
    if (h.n > kMaxN || h.sigma > 256 || h.code_bits > 8 || h.total_bytes < kHeaderBytes)
      return fail(CSFM_ERR_FORMAT, "blob header: n / sigma / code_bits out of range");
    const uint64_t lines_end = h.off_levels + (uint64_t)h.levels * h.level_stride;
    if (h.level_stride > h.total_bytes || lines_end > h.total_bytes || h.nsamp > h.total_bytes / 4)
      return fail(CSFM_ERR_FORMAT, "blob header: section sizes overflow the blob");
    if (h.C[0] != 0 || h.C[256] != h.n) return fail(CSFM_ERR_FORMAT, "blob header: C array does not span [0, n]");
    uint32_t present = 0;
    for (int c = 0; c < 256; ++c) {
      if (h.C[c + 1] < h.C[c]) return fail(CSFM_ERR_FORMAT, "blob header: C array not monotone");
      if (h.C[c + 1] != h.C[c]) {
        ++present;
        if (nib && (h.code_of_byte[c] >> 4) >= (h.levels == 2 ? 16u : 1u))
          return fail(CSFM_ERR_FORMAT, "blob header: compact code does not fit the level count");
        if (dna && h.code_of_byte[c] > kSpecialCode) return fail(CSFM_ERR_FORMAT, "blob header: code beyond the two-bit alphabet");
        if (dna && h.code_of_byte[c] == kSpecialCode && (h.C[c + 1] - h.C[c] != 1 || h.special_byte != (uint32_t)c || h.special_row >= h.n))
          return fail(CSFM_ERR_FORMAT, "blob header: inconsistent single-occurrence symbol");
        if (!nib && !dna && h.levels < 8 && (h.code_of_byte[c] >> h.levels) != 0)
          return fail(CSFM_ERR_FORMAT, "blob header: compact code does not fit the level count");
      }
    }
    if (present != h.sigma) return fail(CSFM_ERR_FORMAT, "blob header: sigma differs from the C array");
    for (int g = 0; g < 16; ++g)
      if (h.start1[g] > h.n || (g && h.start1[g] < h.start1[g - 1]))
        return fail(CSFM_ERR_FORMAT, "blob header: start1 not monotone within [0, n]");
    if (dna && (h.off_text || h.sigma > 5)) return fail(CSFM_ERR_FORMAT, "blob header: layout 3 with text sections or more than five symbols");
    if (!nib && !dna)
      for (uint32_t l = 0; l < h.levels; ++l)
        if (h.zeros[l] > h.n) return fail(CSFM_ERR_FORMAT, "blob header: zeros[] beyond n");
    if (h.kmer_k && h.kmer_radix != (dna ? std::min<uint32_t>(h.sigma, 4u) : h.sigma) && (dna || h.kmer_radix != 256))
      return fail(CSFM_ERR_FORMAT, "blob header: k-mer radix differs from the alphabet");
    if (h.off_text && h.dense_shift != 0) return fail(CSFM_ERR_FORMAT, "blob header: text sections need the full suffix array");
    if (h.verify_min > 64) return fail(CSFM_ERR_FORMAT, "blob header: verify_min out of range");
  }
  auto* idx = new (std::nothrow) csfm_index();
  if (!idx) return fail(CSFM_ERR_NOMEM, "host allocation failed");
  idx->device = device;
  idx->d_blob = static_cast<uint8_t*>(d_blob);
  idx->blob_bytes = h.total_bytes;
  idx->owns_blob = take_ownership != 0;
  idx->h = h;
  rc = index_finish_handle(idx);
  if (rc) {
    idx->owns_blob = false;
    csfm_destroy(idx);
    return rc;
  }
  *out = idx;
  return CSFM_OK;
}

// A second handle over the SAME device blob (no copy): its own streams, workspaces, statistics and mutex, so that
// another host thread can query the index concurrently. The alias borrows the blob: destroy it before the handle
// that owns the blob.
int csfm_alias(const csfm_index* idx, csfm_index** out) {
  if (!idx || !out) return fail(CSFM_ERR_INVALID, "null argument");
  const int rc = csfm_attach_blob(idx->d_blob, idx->blob_bytes, idx->device, 0, out);
  if (rc == CSFM_OK) {
    (*out)->instr_mask = 0;
    (*out)->tma_staging = idx->tma_staging;
    (*out)->no_two_pass = idx->no_two_pass;
    (*out)->no_sa_locate = idx->no_sa_locate;
    (*out)->count3_lanes = idx->count3_lanes;
    (*out)->walk3_lanes = idx->walk3_lanes;
    (*out)->view = idx->view;  // the same experiment knobs as the handle it aliases
  }
  return rc;
}

int csfm_replicate(const csfm_index* idx, int device, csfm_index** out) {
  if (!idx || !out) return fail(CSFM_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  if (device != idx->device) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, idx->device) == cudaSuccess && can) {
      const cudaError_t pe = cudaDeviceEnablePeerAccess(idx->device, 0);
      if (pe != cudaSuccess) (void)cudaGetLastError();  // already enabled, or not possible: the copy below still works
    }
  }
  void* d = nullptr;
  CSFM_CUDA(cudaMalloc(&d, idx->blob_bytes));
  const cudaError_t e = device == idx->device ? cudaMemcpy(d, idx->d_blob, idx->blob_bytes, cudaMemcpyDeviceToDevice)
                                              : cudaMemcpyPeer(d, device, idx->d_blob, idx->device, idx->blob_bytes);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(CSFM_ERR_CUDA, std::string("replicate: ") + cudaGetErrorString(e));
  }
  rc = csfm_attach_blob(d, idx->blob_bytes, device, 1, out);
  if (rc) cudaFree(d);
  return rc;
}

int csfm_blob_to_host(const csfm_index* idx, void* out, uint64_t bytes) {
  if (!idx || !out) return fail(CSFM_ERR_INVALID, "null argument");
  if (bytes < idx->blob_bytes) return fail(CSFM_ERR_CAPACITY, "host buffer smaller than the blob");
  DeviceGuard g(idx->device);
  CSFM_CUDA(cudaMemcpy(out, idx->d_blob, idx->blob_bytes, cudaMemcpyDeviceToHost));
  return CSFM_OK;
}

int csfm_from_host_blob(const void* blob, uint64_t bytes, int device, csfm_index** out) {
  if (!out || !blob) return fail(CSFM_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  if (bytes < kHeaderBytes) return fail(CSFM_ERR_FORMAT, "blob smaller than its header");
  DeviceGuard g(device);
  if (!g.ok) return fail(CSFM_ERR_CUDA, "cudaSetDevice failed");
  void* d = nullptr;
  CSFM_CUDA(cudaMalloc(&d, bytes));
  cudaError_t e = cudaMemcpy(d, blob, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(CSFM_ERR_CUDA, cudaGetErrorString(e));
  }
  rc = csfm_attach_blob(d, bytes, device, 1, out);
  if (rc) cudaFree(d);
  return rc;
}

// ---- queries ----------------------------------------------------------------------------

static void begin_call(csfm_index* idx) {
  std::memset(&idx->stats, 0, sizeof idx->stats);
  std::memset(idx->h_pinned, 0, 128);
}

static int end_call(csfm_index* idx, cudaStream_t stream, bool locate) {
  // caller has synchronised `stream`
  const unsigned long long* hp = static_cast<const unsigned long long*>(idx->h_pinned);
  if (idx->instr_mask & 1u) {
    if (locate) {
      idx->stats.lf_steps = hp[9];
    } else {
      idx->stats.search_steps = hp[8];
      idx->stats.table_lookups = (uint32_t)hp[10];
      idx->stats.text_checks = (uint32_t)hp[11];
      idx->stats.half_steps = (uint32_t)hp[12];
      idx->stats.line_fetches = hp[13];
    }
  }
  if (idx->instr_mask & 2u) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, idx->ev0, idx->ev1) == cudaSuccess) idx->stats.kernel_ms = ms;
  }
  (void)stream;
  return CSFM_OK;
}

// ---- single-query path -------------------------------------------------------------------------
// One launch, no copies: the pattern rides in the kernel parameters, the kernel writes count and interval
// to mapped pinned memory and then the sequence number this thread spins on. Caller holds idx->mu.
namespace {
constexpr size_t kSingleResultAt = 512;  // byte offset of the SingleResult inside h_pinned

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

bool single_path_ok(const csfm_index* idx, uint64_t len) {
  return idx->view.layout == kLayoutNibble128 && len <= kSingleMax && idx->instr_mask == 0 && idx->d_pinned != nullptr;
}

int count_single(csfm_index* idx, const uint8_t* bytes, uint32_t len, uint64_t* count, uint64_t* sp, uint64_t* ep) {
  SingleQuery q;
  const BlobHeader& h = idx->h;
  for (uint32_t i = 0; i < len; ++i) {  // argument marshalling: the per-character table entries ride in the parameters
    const uint32_t b = bytes[i], code = h.code_of_byte[b];
    q.ent[i] = make_uint4(h.base_by_byte[b], code | (h.C[b + 1] == h.C[b] ? 0x80000000u : 0u), h.start1[(code >> 4) & 15u], h.C[b]);
  }
  q.len = len;
  q.c_after_last = len ? h.C[(uint32_t)bytes[len - 1] + 1] : 0u;
  q.pad = 0;
  q.seq = ++idx->single_seq;
  if (q.seq == 0) q.seq = ++idx->single_seq;  // 0 is the value of a fresh result slot
  auto* d_res = reinterpret_cast<SingleResult*>(static_cast<uint8_t*>(idx->d_pinned) + kSingleResultAt);
  volatile SingleResult* res = reinterpret_cast<volatile SingleResult*>(static_cast<uint8_t*>(idx->h_pinned) + kSingleResultAt);
  launch_count_single2(idx->view, q, d_res, idx->stream);
  CSFM_CUDA(cudaGetLastError());
  for (uint32_t spins = 1; res->seq != q.seq; ++spins) {
    cpu_relax();
    if ((spins & 0x3FFFu) == 0) {  // a failed launch never writes the flag: ask the stream now and then
      const cudaError_t e = cudaStreamQuery(idx->stream);
      if (e != cudaErrorNotReady && e != cudaSuccess) return fail(CSFM_ERR_CUDA, std::string("single-query kernel: ") + cudaGetErrorString(e));
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  // result and sequence number were written by ONE 16-byte store (one cache line, one transaction)
  *count = res->count;
  if (sp) *sp = res->sp;
  if (ep) *ep = res->ep;
  idx->stats.kernel_launches += 1;
  idx->stats.h2d_bytes = len;  // as kernel parameters
  idx->stats.d2h_bytes = sizeof(SingleResult);
  return CSFM_OK;
}
}  // namespace

int csfm_count_batch_device(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat,
                            uint64_t* d_counts, uint64_t* d_sp_ep, void* stream) {
  if (!idx || (npat && (!d_offs || !d_counts))) return fail(CSFM_ERR_INVALID, "null argument");
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);  // the per-call stats and the cursor ring are shared by all callers of a handle
  begin_call(idx);
  return count_device(idx, d_bytes, d_offs, npat, d_counts, d_sp_ep, nullptr, nullptr, 0,
                      static_cast<cudaStream_t>(stream));
}

int csfm_count_batch(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs, uint64_t npat, uint64_t* counts,
                     uint64_t* sp_ep) {
  if (!idx || (npat && (!offs || !counts))) return fail(CSFM_ERR_INVALID, "null argument");
  if (npat == 0) return CSFM_OK;
  const uint64_t nbytes = offs[npat];
  if (nbytes && !bytes) return fail(CSFM_ERR_INVALID, "bytes is null");
  if (!offsets_monotone(offs, npat)) return fail(CSFM_ERR_INVALID, "pattern offsets must not decrease");
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  if (npat == 1 && offs[1] >= offs[0] && single_path_ok(idx, offs[1] - offs[0])) {  // cs::FMIndex::count(pattern)
    std::memset(&idx->stats, 0, sizeof idx->stats);
    uint64_t sp = 0, ep = 0;
    const int rc1 = count_single(idx, bytes + offs[0], (uint32_t)(offs[1] - offs[0]), counts, &sp, &ep);
    if (rc1 == CSFM_OK && sp_ep) { sp_ep[0] = sp; sp_ep[1] = ep; }
    return rc1;
  }
  begin_call(idx);
  cudaStream_t st = idx->stream;
  // input staging: [offs (npat+1) u64][bytes]
  const size_t offs_bytes = (npat + 1) * 8;
  int rc = idx->ws_in.ensure(offs_bytes + nbytes + 512);
  if (rc) return rc;
  const size_t out_bytes = npat * 8 * (sp_ep ? 3 : 1);
  rc = idx->ws_out.ensure(out_bytes);
  if (rc) return rc;
  uint64_t* d_offs = idx->ws_in.as<uint64_t>();
  uint8_t* d_bytes = idx->ws_in.as<uint8_t>() + ((offs_bytes + 255) & ~(size_t)255);  // aligned: TMA / cp.async staging
  uint64_t* d_counts = idx->ws_out.as<uint64_t>();
  uint64_t* d_sp_ep = sp_ep ? d_counts + npat : nullptr;
  idx->stats.h2d_bytes = offs_bytes + nbytes;
  idx->stats.d2h_bytes = out_bytes;

  // Large batches are pipelined: slices on alternating streams, so that the host->device copy of
  // slice i+1 and the device->host copy of slice i-1 overlap the kernel of slice i, and the
  // persistent CTAs of consecutive slices hand the SMs over without a gap. Offsets stay absolute,
  // so a slice's bytes land at the same positions of the device buffer as in one big copy.
  constexpr uint64_t kMinSlice = 1ull << 16;
  const uint64_t nslices = (idx->instr_mask == 0 && npat >= 4 * kMinSlice) ? 4 : 1;
  if (nslices == 1) {
    CSFM_CUDA(cudaMemcpyAsync(d_offs, offs, offs_bytes, cudaMemcpyHostToDevice, st));
    if (nbytes) CSFM_CUDA(cudaMemcpyAsync(d_bytes, bytes, nbytes, cudaMemcpyHostToDevice, st));
    rc = count_device(idx, d_bytes, d_offs, npat, d_counts, d_sp_ep, nullptr, nullptr, 0, st);
    if (rc) return rc;
    CSFM_CUDA(cudaMemcpyAsync(counts, d_counts, npat * 8, cudaMemcpyDeviceToHost, st));
    if (sp_ep) CSFM_CUDA(cudaMemcpyAsync(sp_ep, d_sp_ep, npat * 16, cudaMemcpyDeviceToHost, st));
    CSFM_CUDA(cudaStreamSynchronize(st));
    return end_call(idx, st, false);
  }
  for (int s = 0; s < 2; ++s)
    if (!idx->aux_stream[s]) CSFM_CUDA(cudaStreamCreateWithFlags(&idx->aux_stream[s], cudaStreamNonBlocking));
  for (uint64_t s = 0; s < nslices; ++s) {
    const uint64_t lo = npat * s / nslices, hi = npat * (s + 1) / nslices;
    cudaStream_t ss = idx->aux_stream[s & 1];
    // offsets lo..hi inclusive (the shared boundary entry is copied twice with the same value)
    CSFM_CUDA(cudaMemcpyAsync(d_offs + lo, offs + lo, (hi - lo + 1) * 8, cudaMemcpyHostToDevice, ss));
    const uint64_t b0 = offs[lo], b1 = offs[hi];
    if (b1 > b0) CSFM_CUDA(cudaMemcpyAsync(d_bytes + b0, bytes + b0, b1 - b0, cudaMemcpyHostToDevice, ss));
    rc = count_device(idx, d_bytes, d_offs + lo, hi - lo, d_counts + lo, d_sp_ep ? d_sp_ep + 2 * lo : nullptr, nullptr,
                      nullptr, 0, ss);
    if (rc) return rc;
    CSFM_CUDA(cudaMemcpyAsync(counts + lo, d_counts + lo, (hi - lo) * 8, cudaMemcpyDeviceToHost, ss));
    if (sp_ep) CSFM_CUDA(cudaMemcpyAsync(sp_ep + 2 * lo, d_sp_ep + 2 * lo, (hi - lo) * 16, cudaMemcpyDeviceToHost, ss));
  }
  CSFM_CUDA(cudaStreamSynchronize(idx->aux_stream[0]));
  CSFM_CUDA(cudaStreamSynchronize(idx->aux_stream[1]));
  return end_call(idx, st, false);
}

namespace {
__global__ void widen_offsets_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
__global__ void narrow_counts_kernel(const uint64_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (uint32_t)in[i];
}

// packed wire codes -> pattern bytes: symbol i of the batch sits in bits [i * bits, (i + 1) * bits) of `packed`
struct WireTable {
  uint8_t byte_of_wire[256];
};
__global__ void unpack_codes_kernel(const uint8_t* __restrict__ packed, uint64_t packed_bytes, uint64_t nsyms, uint32_t bits,
                                    const __grid_constant__ WireTable wt, uint8_t* __restrict__ out) {
  const uint32_t mask = (1u << bits) - 1u;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nsyms; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t bit = i * bits, at = bit >> 3;
    uint32_t w = packed[at];
    if (at + 1 < packed_bytes) w |= (uint32_t)packed[at + 1] << 8;
    out[i] = wt.byte_of_wire[(w >> (bit & 7)) & mask];
  }
}

// wire codes of an index: the bytes that occur, in byte order; a code beyond them decodes to a byte that does not
// occur (so the pattern counts 0), or — with all 256 bytes present — cannot be formed at all
void wire_codes(const csfm_index* idx, uint8_t code_of_byte[256], uint8_t byte_of_wire[256], uint32_t* bits) {
  uint32_t sigma = 0;
  int absent = -1;
  for (int c = 0; c < 256; ++c) {
    if (idx->h.C[c + 1] != idx->h.C[c]) {
      if (code_of_byte) code_of_byte[c] = (uint8_t)sigma;
      byte_of_wire[sigma++] = (uint8_t)c;
    } else {
      if (code_of_byte) code_of_byte[c] = 255;
      if (absent < 0) absent = c;
    }
  }
  for (uint32_t k = sigma; k < 256; ++k) byte_of_wire[k] = (uint8_t)(absent < 0 ? 0 : absent);
  uint32_t b = 1;
  while ((1u << b) < sigma) ++b;
  *bits = b;
}

// How a streaming batch describes where its patterns start.
enum class PatternIndex { kOffsets64, kOffsets32, kLengths8, kPackedLengths8 };

// The three submit entry points differ only in how the pattern boundaries arrive and how wide the
// counts leave: one slot, one stream, H2D copies -> (offsets rebuilt on the device) -> count kernel ->
// (counts narrowed) -> D2H copy. Exactly one of counts64 / counts32 is set.
int submit_batch(csfm_index* idx, const uint8_t* bytes, uint64_t nbytes, PatternIndex kind, const void* index,
                 uint64_t npat, uint64_t* counts64, uint64_t* sp_ep, uint32_t* counts32, uint64_t* ticket) {
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  const uint64_t t = idx->next_ticket++;
  csfm_index::AsyncSlot& sl = idx->async_slot[t % CSFM_ASYNC_SLOTS];
  if (!sl.stream) CSFM_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  if (sl.ticket) CSFM_CUDA(cudaStreamSynchronize(sl.stream));  // slot still busy with an older batch
  sl.ticket = t;
  *ticket = t;
  if (npat == 0) return CSFM_OK;
  // input slot: [offs (npat+1) u64][bytes, 256-byte aligned: TMA / cp.async staging][compact index as copied, 256-byte aligned]
  const size_t offs_bytes = (npat + 1) * 8;
  const size_t bytes_at = (offs_bytes + 255) & ~(size_t)255;
  const size_t compact_at = (bytes_at + nbytes + 255) & ~(size_t)255;
  const bool packed = kind == PatternIndex::kPackedLengths8;
  uint32_t wire_bits = 8;
  WireTable wt;
  if (packed) wire_codes(idx, nullptr, wt.byte_of_wire, &wire_bits);
  const uint64_t packed_bytes = packed ? (nbytes * wire_bits + 7) / 8 : 0;  // nbytes = symbols of the batch
  const size_t compact_bytes = kind == PatternIndex::kOffsets32 ? (npat + 1) * 4 : (kind == PatternIndex::kLengths8 || packed) ? npat + 1 : 0;
  const size_t packed_at = (compact_at + compact_bytes + 255) & ~(size_t)255;
  int rc = sl.in.ensure(packed_at + packed_bytes + 512);
  if (rc) return rc;
  rc = sl.out.ensure(npat * 8 * (sp_ep ? 3 : 1) + (counts32 ? npat * 4 : 0));  // [counts u64][sp_ep][counts u32]
  if (rc) return rc;
  uint64_t* d_offs = sl.in.as<uint64_t>();
  uint8_t* d_bytes = sl.in.as<uint8_t>() + bytes_at;
  uint8_t* d_compact = sl.in.as<uint8_t>() + compact_at;
  uint64_t* d_counts = sl.out.as<uint64_t>();
  uint64_t* d_sp_ep = sp_ep ? d_counts + npat : nullptr;
  uint32_t* d_counts32 = reinterpret_cast<uint32_t*>(d_counts + npat * (sp_ep ? 3 : 1));
  const int grid = (int)std::min<uint64_t>((npat + 1 + 255) / 256, (uint64_t)idx->num_sms * 8);
  // all host->device copies first: a kernel between two copies would make the second copy wait until that
  // kernel has found room beside the persistent count kernel of another slot
  switch (kind) {
    case PatternIndex::kOffsets64:
      CSFM_CUDA(cudaMemcpyAsync(d_offs, index, offs_bytes, cudaMemcpyHostToDevice, sl.stream));
      idx->stats.h2d_bytes = offs_bytes + nbytes;
      break;
    case PatternIndex::kOffsets32:
      CSFM_CUDA(cudaMemcpyAsync(d_compact, index, (npat + 1) * 4, cudaMemcpyHostToDevice, sl.stream));
      idx->stats.h2d_bytes = (npat + 1) * 4 + nbytes;
      break;
    case PatternIndex::kLengths8:
    case PatternIndex::kPackedLengths8:
      CSFM_CUDA(cudaMemcpyAsync(d_compact, index, npat, cudaMemcpyHostToDevice, sl.stream));
      CSFM_CUDA(cudaMemsetAsync(d_compact + npat, 0, 1, sl.stream));  // so that offs[npat] comes out of the same scan
      idx->stats.h2d_bytes = npat + (packed ? packed_bytes : nbytes);
      break;
  }
  if (packed) {
    uint8_t* d_packed = sl.in.as<uint8_t>() + packed_at;
    if (packed_bytes) {
      CSFM_CUDA(cudaMemcpyAsync(d_packed, bytes, packed_bytes, cudaMemcpyHostToDevice, sl.stream));
      const int g = (int)std::min<uint64_t>((nbytes + 255) / 256, (uint64_t)idx->num_sms * 16);
      unpack_codes_kernel<<<g, 256, 0, sl.stream>>>(d_packed, packed_bytes, nbytes, wire_bits, wt, d_bytes);
      CSFM_CUDA(cudaGetLastError());
    }
  } else if (nbytes) {
    CSFM_CUDA(cudaMemcpyAsync(d_bytes, bytes, nbytes, cudaMemcpyHostToDevice, sl.stream));
  }
  if (kind == PatternIndex::kOffsets32) {
    widen_offsets_kernel<<<grid, 256, 0, sl.stream>>>(reinterpret_cast<const uint32_t*>(d_compact), d_offs, npat + 1);
    CSFM_CUDA(cudaGetLastError());
  } else if (kind == PatternIndex::kLengths8 || packed) {
    rc = offsets_from_lengths8(d_compact, npat + 1, d_offs, sl.scan, sl.stream);
    if (rc) return rc;
  }
  const uint32_t saved = idx->instr_mask;
  idx->instr_mask = 0;  // per-call instrumentation is a property of the synchronous entry points
  rc = count_device(idx, d_bytes, d_offs, npat, d_counts, d_sp_ep, nullptr, nullptr, 0, sl.stream);
  idx->instr_mask = saved;
  if (rc) return rc;
  if (counts32) {
    narrow_counts_kernel<<<grid, 256, 0, sl.stream>>>(d_counts, d_counts32, npat);
    CSFM_CUDA(cudaGetLastError());
    CSFM_CUDA(cudaMemcpyAsync(counts32, d_counts32, npat * 4, cudaMemcpyDeviceToHost, sl.stream));
    idx->stats.d2h_bytes = npat * 4;
  } else {
    CSFM_CUDA(cudaMemcpyAsync(counts64, d_counts, npat * 8, cudaMemcpyDeviceToHost, sl.stream));
    if (sp_ep) CSFM_CUDA(cudaMemcpyAsync(sp_ep, d_sp_ep, npat * 16, cudaMemcpyDeviceToHost, sl.stream));
    idx->stats.d2h_bytes = npat * 8 * (sp_ep ? 3 : 1);
  }
  return CSFM_OK;
}
}  // namespace

int csfm_count_batch_submit(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs, uint64_t npat, uint64_t* counts,
                            uint64_t* sp_ep, uint64_t* ticket) {
  if (!idx || !ticket || (npat && (!offs || !counts))) return fail(CSFM_ERR_INVALID, "null argument");
  const uint64_t nbytes = npat ? offs[npat] : 0;
  if (nbytes && !bytes) return fail(CSFM_ERR_INVALID, "bytes is null");
  if (npat && !offsets_monotone(offs, npat)) return fail(CSFM_ERR_INVALID, "pattern offsets must not decrease");
  return submit_batch(idx, bytes, nbytes, PatternIndex::kOffsets64, offs, npat, counts, sp_ep, nullptr, ticket);
}

int csfm_count_batch_submit32(csfm_index* idx, const uint8_t* bytes, const uint32_t* offs32, uint64_t npat,
                              uint32_t* counts32, uint64_t* ticket) {
  if (!idx || !ticket || (npat && (!offs32 || !counts32))) return fail(CSFM_ERR_INVALID, "null argument");
  const uint64_t nbytes = npat ? offs32[npat] : 0;
  if (nbytes && !bytes) return fail(CSFM_ERR_INVALID, "bytes is null");
  if (npat && !offsets_monotone(offs32, npat)) return fail(CSFM_ERR_INVALID, "pattern offsets must not decrease");
  return submit_batch(idx, bytes, nbytes, PatternIndex::kOffsets32, offs32, npat, nullptr, nullptr, counts32, ticket);
}

int csfm_count_batch_submit_len8(csfm_index* idx, const uint8_t* bytes, uint64_t nbytes, const uint8_t* lens8,
                                 uint64_t npat, uint32_t* counts32, uint64_t* ticket) {
  if (!idx || !ticket || (npat && (!lens8 || !counts32))) return fail(CSFM_ERR_INVALID, "null argument");
  if (nbytes && !bytes) return fail(CSFM_ERR_INVALID, "bytes is null");
  uint64_t sum = 0;  // the kernel trusts the rebuilt offsets: they must stay inside the staged bytes
  for (uint64_t i = 0; i < npat; ++i) sum += lens8[i];
  if (sum != nbytes) return fail(CSFM_ERR_INVALID, "nbytes differs from the sum of the pattern lengths");
  return submit_batch(idx, bytes, nbytes, PatternIndex::kLengths8, lens8, npat, nullptr, nullptr, counts32, ticket);
}

int csfm_pattern_codes(const csfm_index* idx, uint8_t code_of_byte[256], uint32_t* bits) {
  if (!idx || !code_of_byte || !bits) return fail(CSFM_ERR_INVALID, "null argument");
  uint8_t byte_of_wire[256];
  wire_codes(idx, code_of_byte, byte_of_wire, bits);
  return CSFM_OK;
}

int csfm_count_batch_submit_packed(csfm_index* idx, const uint8_t* packed, uint64_t nsyms, const uint8_t* lens8, uint64_t npat,
                                   uint32_t* counts32, uint64_t* ticket) {
  if (!idx || !ticket || (npat && (!lens8 || !counts32))) return fail(CSFM_ERR_INVALID, "null argument");
  if (nsyms && !packed) return fail(CSFM_ERR_INVALID, "packed is null");
  uint64_t sum = 0;  // the kernel trusts the rebuilt offsets: they must stay inside the unpacked symbols
  for (uint64_t i = 0; i < npat; ++i) sum += lens8[i];
  if (sum != nsyms) return fail(CSFM_ERR_INVALID, "nsyms differs from the sum of the pattern lengths");
  return submit_batch(idx, packed, nsyms, PatternIndex::kPackedLengths8, lens8, npat, nullptr, nullptr, counts32, ticket);
}

int csfm_count_batch_wait(csfm_index* idx, uint64_t ticket) {
  if (!idx || ticket == 0) return fail(CSFM_ERR_INVALID, "bad ticket");
  DeviceGuard g(idx->device);
  cudaStream_t st = nullptr;
  {
    std::lock_guard<std::mutex> lk(idx->mu);
    csfm_index::AsyncSlot& sl = idx->async_slot[ticket % CSFM_ASYNC_SLOTS];
    if (sl.ticket != ticket) return CSFM_OK;  // already waited on (or recycled, which waited on it)
    st = sl.stream;
  }
  CSFM_CUDA(cudaStreamSynchronize(st));
  std::lock_guard<std::mutex> lk(idx->mu);
  csfm_index::AsyncSlot& sl = idx->async_slot[ticket % CSFM_ASYNC_SLOTS];
  if (sl.ticket == ticket) sl.ticket = 0;
  return CSFM_OK;
}

int csfm_locate_batch_device(csfm_index* idx, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t npat,
                             uint64_t limit, uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t cap,
                             int32_t* d_status, uint64_t* total, void* stream) {
  if (!idx || !total || !d_out_offs || (npat && !d_offs)) return fail(CSFM_ERR_INVALID, "null argument");
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  begin_call(idx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = locate_plan(idx, d_bytes, d_offs, npat, limit, d_out_offs, d_status, total, st);
  if (rc) return rc;
  if (d_out_pos == nullptr && cap == 0) return CSFM_OK;  // sizing call
  if (cap < *total) return fail(CSFM_ERR_CAPACITY, "locate output buffer too small");
  rc = locate_expand(idx, npat, d_out_offs, d_out_pos, *total, st);
  if (rc) return rc;
  return locate_walk(idx, npat, d_out_offs, d_out_pos, 0, *total, d_status, st);
}

int csfm_locate_batch(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs, uint64_t npat, uint64_t limit,
                      uint64_t* out_offs, uint64_t* out_pos, uint64_t cap, int32_t* status, uint64_t* total) {
  if (!idx || !total || !out_offs || (npat && !offs)) return fail(CSFM_ERR_INVALID, "null argument");
  *total = 0;
  if (npat == 0) {
    out_offs[0] = 0;
    return CSFM_OK;
  }
  const uint64_t nbytes = offs[npat];
  if (nbytes && !bytes) return fail(CSFM_ERR_INVALID, "bytes is null");
  if (!offsets_monotone(offs, npat)) return fail(CSFM_ERR_INVALID, "pattern offsets must not decrease");
  DeviceGuard g(idx->device);
  std::lock_guard<std::mutex> lk(idx->mu);
  if (npat == 1 && offs[1] >= offs[0] && single_path_ok(idx, offs[1] - offs[0]) && limit <= 0xFFFFFFFFull) {
    // cs::FMIndex::locate(pattern): the interval from the single-query kernel (remembered between the
    // sizing call and the fill call), then one walk launch over the rows sp .. sp + min(count, limit)
    std::memset(&idx->stats, 0, sizeof idx->stats);
    const uint8_t* pat = bytes + offs[0];
    const uint32_t len = (uint32_t)(offs[1] - offs[0]);
    csfm_index::SingleCache& sc = idx->single_cache;
    if (!(sc.valid && sc.len == len && std::memcmp(sc.bytes, pat, len) == 0)) {
      uint64_t cnt = 0, sp = 0;
      const int rc1 = count_single(idx, pat, len, &cnt, &sp, nullptr);
      if (rc1) return rc1;
      sc.valid = true;
      sc.len = len;
      if (len) std::memcpy(sc.bytes, pat, len);
      sc.sp = sp;
      sc.count = len ? cnt : 0;  // locate("") is empty (fm_index.cpp:109)
    }
    const uint64_t tot = std::min<uint64_t>(sc.count, limit);
    out_offs[0] = 0;
    out_offs[1] = tot;
    *total = tot;
    if (status) status[0] = 0;
    const bool sizing1 = (out_pos == nullptr && cap == 0);
    if (sizing1) return CSFM_OK;
    if (cap < tot) return fail(CSFM_ERR_CAPACITY, "locate output buffer too small");
    if (tot == 0) return CSFM_OK;
    int rc1 = idx->ws_pos.ensure((tot + 1) * 8);
    if (rc1) return rc1;
    uint64_t* d_pos = idx->ws_pos.as<uint64_t>();
    int32_t* d_stat = reinterpret_cast<int32_t*>(d_pos + tot);
    cudaStream_t s1 = idx->stream;
    CSFM_CUDA(cudaMemsetAsync(d_stat, 0, 8, s1));
    rc1 = locate_walk(idx, 1, nullptr, d_pos, 0, tot, d_stat, s1, (int64_t)sc.sp);
    if (rc1) return rc1;
    CSFM_CUDA(cudaMemcpyAsync(out_pos, d_pos, tot * 8, cudaMemcpyDeviceToHost, s1));
    int32_t st1 = 0;
    CSFM_CUDA(cudaMemcpyAsync(&st1, d_stat, 4, cudaMemcpyDeviceToHost, s1));
    CSFM_CUDA(cudaStreamSynchronize(s1));
    if (status) status[0] = st1;
    idx->stats.d2h_bytes = tot * 8 + 4;
    return CSFM_OK;
  }
  begin_call(idx);
  cudaStream_t st = idx->stream;
  const size_t offs_bytes = (npat + 1) * 8;
  const size_t status_bytes = (npat * 4 + 7) / 8 * 8;
  // ws_in: [offs][bytes] ; ws_out: [out_offs (npat+1) u64][status npat i32] ; ws_pos: positions
  int rc = idx->ws_in.ensure(offs_bytes + nbytes + 512);
  if (rc) return rc;
  rc = idx->ws_out.ensure(offs_bytes + status_bytes);
  if (rc) return rc;
  uint64_t* d_offs = idx->ws_in.as<uint64_t>();
  uint8_t* d_bytes = idx->ws_in.as<uint8_t>() + ((offs_bytes + 255) & ~(size_t)255);  // aligned: TMA / cp.async staging
  uint64_t* d_out_offs = idx->ws_out.as<uint64_t>();
  int32_t* d_status = reinterpret_cast<int32_t*>(idx->ws_out.as<uint8_t>() + offs_bytes);
  CSFM_CUDA(cudaMemcpyAsync(d_offs, offs, offs_bytes, cudaMemcpyHostToDevice, st));
  if (nbytes) CSFM_CUDA(cudaMemcpyAsync(d_bytes, bytes, nbytes, cudaMemcpyHostToDevice, st));
  idx->stats.h2d_bytes = offs_bytes + nbytes;

  uint64_t tot = 0;
  rc = locate_plan(idx, d_bytes, d_offs, npat, limit, d_out_offs, d_status, &tot, st);
  if (rc) return rc;
  *total = tot;
  CSFM_CUDA(cudaMemcpyAsync(out_offs, d_out_offs, offs_bytes, cudaMemcpyDeviceToHost, st));
  idx->stats.d2h_bytes = offs_bytes;
  const bool sizing = (out_pos == nullptr && cap == 0);
  if (sizing || cap < tot) {
    if (status) CSFM_CUDA(cudaMemcpyAsync(status, d_status, npat * 4, cudaMemcpyDeviceToHost, st));
    CSFM_CUDA(cudaStreamSynchronize(st));
    return sizing ? CSFM_OK : fail(CSFM_ERR_CAPACITY, "locate output buffer too small");
  }
  if (tot) {
    rc = idx->ws_pos.ensure(tot * 8);
    if (rc) return rc;
    uint64_t* d_pos = idx->ws_pos.as<uint64_t>();
    rc = locate_expand(idx, npat, d_out_offs, d_pos, tot, st);
    if (rc) return rc;
    // Large results leave in slices: the device->host copy of a slice (8 bytes per occurrence) runs on a
    // second stream while the next slice is being walked.
    const uint64_t nslices = (idx->instr_mask == 0 && tot >= (1ull << 22)) ? 4 : 1;
    if (nslices > 1 && !idx->aux_stream[0]) CSFM_CUDA(cudaStreamCreateWithFlags(&idx->aux_stream[0], cudaStreamNonBlocking));
    for (uint64_t s = 0; s < nslices; ++s) {
      const uint64_t lo = tot * s / nslices, hi = tot * (s + 1) / nslices;
      rc = locate_walk(idx, npat, d_out_offs, d_pos, lo, hi - lo, d_status, st);
      if (rc) return rc;
      cudaStream_t cs = st;
      if (nslices > 1) {
        if (!idx->ev_slice[s]) CSFM_CUDA(cudaEventCreateWithFlags(&idx->ev_slice[s], cudaEventDisableTiming));
        CSFM_CUDA(cudaEventRecord(idx->ev_slice[s], st));
        cs = idx->aux_stream[0];
        CSFM_CUDA(cudaStreamWaitEvent(cs, idx->ev_slice[s], 0));
      }
      CSFM_CUDA(cudaMemcpyAsync(out_pos + lo, d_pos + lo, (hi - lo) * 8, cudaMemcpyDeviceToHost, cs));
    }
    if (nslices > 1) CSFM_CUDA(cudaStreamSynchronize(idx->aux_stream[0]));
  }
  if (status) CSFM_CUDA(cudaMemcpyAsync(status, d_status, npat * 4, cudaMemcpyDeviceToHost, st));
  CSFM_CUDA(cudaStreamSynchronize(st));
  idx->stats.d2h_bytes += tot * 8 + (status ? npat * 4 : 0);
  return end_call(idx, st, true);
}

int csfm_set_instrumentation(csfm_index* idx, uint32_t mask) {
  if (!idx) return fail(CSFM_ERR_INVALID, "null argument");
  idx->instr_mask = mask;
  return CSFM_OK;
}

int csfm_last_call_stats(const csfm_index* idx, csfm_call_stats* out) {
  if (!idx || !out) return fail(CSFM_ERR_INVALID, "null argument");
  // device-pointer calls are asynchronous: fold in whatever the instrumentation wrote so far
  *out = idx->stats;
  const unsigned long long* hp = static_cast<const unsigned long long*>(idx->h_pinned);
  if (idx->instr_mask & 1u) {
    if (!out->search_steps) out->search_steps = hp[8];
    if (!out->table_lookups) out->table_lookups = (uint32_t)hp[10];
    if (!out->text_checks) out->text_checks = (uint32_t)hp[11];
    if (!out->half_steps) out->half_steps = (uint32_t)hp[12];
    if (!out->line_fetches) out->line_fetches = hp[13];
    if (!out->lf_steps) out->lf_steps = hp[9];
  }
  if ((idx->instr_mask & 2u) && out->kernel_ms == 0.f) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, idx->ev0, idx->ev1) == cudaSuccess) out->kernel_ms = ms;
  }
  return CSFM_OK;
}

int csfm_host_alloc(void** p, uint64_t bytes) {
  if (!p) return fail(CSFM_ERR_INVALID, "null argument");
  *p = nullptr;
  cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? CSFM_ERR_NOMEM : CSFM_ERR_CUDA, cudaGetErrorString(e));
  return CSFM_OK;
}

int csfm_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return CSFM_OK;
}

}  // extern "C"
