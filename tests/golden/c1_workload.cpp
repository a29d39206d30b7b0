// tests/golden/c1_workload.cpp — golden-vector GENERATOR input (not product code).
//
// Regenerates the inputs of the reference's published benchmark (BASELINE.json configs[0]):
//   text     = generate_text_with_patterns(100000) + '$'     (tools/benchmark.cpp:55-76, :276-277)
//   random   = 10 000 substrings of length 5 at dist(rng)     (tools/benchmark.cpp:82-96, seed 42)
//   frequent = 10 fixed bi/trigrams cycled to 10 000           (tools/benchmark.cpp:98-112)
// std::uniform_int_distribution is implementation-defined, so this must be compiled with the
// same libstdc++ as the reference probe in SURVEY §6 (g++ 13.3); make_golden.py then checks the
// reference's total_matches checksums (9 907 582 / 16 309 000 / 163 090) before writing
// c1_workload.npz.
// Output (binary, stdout): u64 n, n text bytes, u64 npat, npat u32 start positions.
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

int main() {
  const size_t length = 100000;
  const char* words[8] = {"banana", "apple", "orange", "grape", "cherry",
                          "the quick brown fox", "jumps over", "lazy dog"};
  std::string text;
  text.reserve(length + 32);
  std::mt19937 rng(12345);
  std::uniform_int_distribution<int> pick(0, 7);
  while (text.size() < length) {
    text += words[pick(rng)];
    text += " ";
  }
  text.resize(length);
  text += "$";

  const size_t npat = 10000, plen = 5;
  std::mt19937 prng(42);
  std::uniform_int_distribution<size_t> dist(0, text.size() - plen - 1);
  std::vector<uint32_t> pos(npat);
  for (size_t i = 0; i < npat; ++i) pos[i] = static_cast<uint32_t>(dist(prng));

  uint64_t n = text.size(), np = npat;
  fwrite(&n, 8, 1, stdout);
  fwrite(text.data(), 1, n, stdout);
  fwrite(&np, 8, 1, stdout);
  fwrite(pos.data(), 4, npat, stdout);
  return 0;
}
