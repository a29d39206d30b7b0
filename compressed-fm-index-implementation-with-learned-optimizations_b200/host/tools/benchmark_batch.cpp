// host/tools/benchmark_batch.cpp — the `--batch` companion of the reference's tools/benchmark.cpp
// (/root/reference/tools/benchmark.cpp:129-224 times count()/locate() one query at a time).
//
// Same three legs, same statistics, on pattern lists read from a file, through the drop-in
// cs::FMIndex (host/src/api/fm_index.hpp):
//   * single-query loops: count(p) / locate(p) per pattern with per-query latency (QPS, p50/p95/p99) —
//     what the reference tool measures, and
//   * with --batch: the whole list in ONE count_batch / locate_batch call (queries/s, occurrences/s).
// Prints one JSON object. Inputs (written by bench.py from tests/golden/c1_workload.npz, or by any caller):
//   text file: raw bytes;  pattern file: u32 count, then per pattern u32 length + bytes (little endian).
//
//   cs_benchmark_batch TEXT COUNT_PATTERNS LOCATE_PATTERNS [--batch] [--repeat R] [--stride S]
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <string_view>
#include <vector>

#include "../src/api/fm_index.hpp"

using Clock = std::chrono::steady_clock;

static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

static std::string slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
  return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

static std::vector<std::string> read_patterns(const char* path) {
  const std::string raw = slurp(path);
  std::vector<std::string> out;
  size_t at = 0;
  auto u32 = [&]() { uint32_t v = 0; if (at + 4 <= raw.size()) std::memcpy(&v, raw.data() + at, 4); at += 4; return v; };
  const uint32_t count = u32();
  out.reserve(count);
  for (uint32_t i = 0; i < count && at <= raw.size(); ++i) {
    const uint32_t len = u32();
    if (at + len > raw.size()) break;
    out.emplace_back(raw.data() + at, len);
    at += len;
  }
  return out;
}

struct Stats { double qps, p50, p95, p99, total_ms; unsigned long long matches; };

template <class F>
static Stats single_loop(const std::vector<std::string>& pats, size_t warmup, F&& query) {
  for (size_t i = 0; i < warmup && !pats.empty(); ++i) (void)query(pats[i % pats.size()]);
  std::vector<double> lat;
  lat.reserve(pats.size());
  Stats s{};
  const auto t0 = Clock::now();
  for (const auto& p : pats) {
    const auto q0 = Clock::now();
    s.matches += query(p);
    lat.push_back(std::chrono::duration<double, std::micro>(Clock::now() - q0).count());
  }
  s.total_ms = ms_since(t0);
  std::sort(lat.begin(), lat.end());
  if (!lat.empty()) {
    s.qps = pats.size() / s.total_ms * 1000.0;
    s.p50 = lat[lat.size() / 2];
    s.p95 = lat[lat.size() * 95 / 100];
    s.p99 = lat[lat.size() * 99 / 100];
  }
  return s;
}

static void print_stats(const char* name, const Stats& s, size_t n, bool last = false) {
  std::printf("  \"%s\": {\"queries\": %zu, \"total_ms\": %.3f, \"qps\": %.1f, \"p50_us\": %.2f, \"p95_us\": %.2f, \"p99_us\": %.2f, "
              "\"total_matches\": %llu}%s\n", name, n, s.total_ms, s.qps, s.p50, s.p95, s.p99, s.matches, last ? "" : ",");
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s TEXT COUNT_PATTERNS LOCATE_PATTERNS [--batch] [--repeat R] [--stride S]\n", argv[0]);
    return 2;
  }
  bool batch = false;
  int repeat = 5;
  cs::BuildParams bp;
  for (int i = 4; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--batch")) batch = true;
    else if (!std::strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = std::max(1, std::atoi(argv[++i]));
    else if (!std::strcmp(argv[i], "--stride") && i + 1 < argc) bp.ssa_stride = (uint32_t)std::atoi(argv[++i]);
  }
  const std::string text = slurp(argv[1]);
  const std::vector<std::string> cpats = read_patterns(argv[2]);
  const std::vector<std::string> lpats = read_patterns(argv[3]);

  (void)cs::FMIndex::build_from_text(text.substr(0, std::min<size_t>(text.size(), 64)), bp);  // CUDA context + module load
  const auto tb = Clock::now();
  cs::FMIndex index = cs::FMIndex::build_from_text(text, bp);
  const double build_ms = ms_since(tb);

  std::printf("{\n  \"text_bytes\": %zu, \"build_ms\": %.3f, \"ssa_stride\": %u,\n", text.size(), build_ms, bp.ssa_stride);
  const Stats c = single_loop(cpats, 100, [&](const std::string& p) { return (unsigned long long)index.count(p); });
  print_stats("count_single", c, cpats.size());
  const Stats l = single_loop(lpats, 10, [&](const std::string& p) { return (unsigned long long)index.locate(p).size(); });
  print_stats("locate_single", l, lpats.size(), !batch);
  if (batch) {
    std::vector<std::string_view> cv(cpats.begin(), cpats.end()), lv(lpats.begin(), lpats.end());
    unsigned long long cm = 0, lm = 0;
    (void)index.count_batch(cv);
    double best_c = 1e30, best_l = 1e30;
    for (int r = 0; r < repeat; ++r) {
      const auto t0 = Clock::now();
      const std::vector<uint64_t> counts = index.count_batch(cv);
      best_c = std::min(best_c, ms_since(t0));
      cm = 0;
      for (uint64_t x : counts) cm += x;
    }
    (void)index.locate_batch(lv);
    for (int r = 0; r < repeat; ++r) {
      const auto t0 = Clock::now();
      const cs::LocateBatch lb = index.locate_batch(lv);
      best_l = std::min(best_l, ms_since(t0));
      lm = lb.positions.size();
    }
    std::printf("  \"count_batch\": {\"queries\": %zu, \"best_ms\": %.4f, \"qps\": %.1f, \"total_matches\": %llu},\n", cv.size(), best_c,
                cv.size() / best_c * 1000.0, cm);
    std::printf("  \"locate_batch\": {\"queries\": %zu, \"best_ms\": %.4f, \"qps\": %.1f, \"occurrences\": %llu, \"occurrences_per_s\": %.1f}\n",
                lv.size(), best_l, lv.size() / best_l * 1000.0, lm, lm / best_l * 1000.0);
  }
  std::printf("}\n");
  return 0;
}
