#pragma once
// host/src/serialization/csidx.hpp — the reference's `.csidx` container
// (/root/reference/src/serialization/serialization.hpp:35-83, serialization.cpp:64-147, :279-335),
// written and read by plain host C++ (no CUDA here).
//
// Layout kept byte for byte: 88-byte IndexHeader {char magic[8]="CSIDX"; u16 version=1; u16 pad;
// u32 flags; u64 text_len; u64 offsets[8]}, sections 8-byte aligned with zero padding, arrays as
// [u64 count][payload]: TEXT, BWT, C_ARRAY (u32), SSA ([u32 stride][pad to 8][u64 count][u32...]),
// WAVELET ([u64 num_levels][3 arrays]), one 4096-aligned opaque layout section, FOOTER
// u64 0x444E4553435300. The reference's writer never terminates when padding is needed
// (serialization.cpp:44-54); its READER works and is the format oracle
// (tests/test_csidx_cpu.py reads our files back through it).
//
// The opaque layout section (the reference's SECTION_VEB_LAYOUT slot, which its own FMIndex never
// reads) carries the device-resident index blob, flagged FLAG_DEVICE_BLOB, so that loading is one
// read + one cudaMemcpy. The WAVELET section is written with num_levels = 0 and empty arrays: the
// reference's two-level directory is not part of this engine.
#include <cstdint>
#include <string>
#include <vector>

namespace cs {

constexpr uint32_t CSIDX_FLAG_DEVICE_BLOB = 1u << 8;  // beyond the reference's bits 0..3 (serialization.hpp:39-45)

struct CsidxSections {
  uint32_t flags = 0;
  bool has_text = false;
  std::string text;
  std::vector<uint8_t> bwt;
  std::vector<uint32_t> c_array;  // FMIndex::C_ has 257 entries (fm_index.cpp:36)
  std::vector<uint32_t> ssa;
  uint32_t ssa_stride = 0;
  bool has_ssa = false;
  std::vector<uint8_t> device_blob;  // opaque; written 4096-aligned
  uint64_t text_len = 0;
};

/// Throws std::runtime_error on I/O failure.
void write_csidx(const std::string& path, const CsidxSections& s);
/// Throws std::runtime_error on I/O failure or a malformed file (magic, version, offsets, footer).
CsidxSections read_csidx(const std::string& path);

}  // namespace cs
