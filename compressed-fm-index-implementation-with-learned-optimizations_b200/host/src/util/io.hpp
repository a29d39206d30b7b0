#pragma once
// host/src/util/io.hpp — whole-file read/write helpers with the reference's names
// (/root/reference/src/util/io.hpp:8-21), used by tools/query_cli.cpp and tools/bench.cpp.
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>

namespace cs {

inline std::string slurp(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open: " + path);
  return {std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()};
}

inline void dump(const std::string& path, const void* data, size_t nbytes) {
  std::ofstream out(path, std::ios::binary | std::ios::trunc);
  if (!out) throw std::runtime_error("cannot write: " + path);
  out.write(static_cast<const char*>(data), static_cast<std::streamsize>(nbytes));
}

inline void dump_str(const std::string& path, const std::string& s) { dump(path, s.data(), s.size()); }

}  // namespace cs
