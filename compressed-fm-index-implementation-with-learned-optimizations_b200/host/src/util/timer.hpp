#pragma once
// host/src/util/timer.hpp — the two helper types the reference's tools use
// (/root/reference/src/util/timer.hpp:9-43): a steady-clock stopwatch and an RAII scope timer
// that reports "[TIMER] name: N ms" on stderr.
#include <chrono>
#include <iostream>
#include <string>
#include <utility>

namespace cs {

class Timer {
public:
  Timer() { reset(); }
  void reset() { t0_ = clock::now(); }
  double elapsed_ms() const { return std::chrono::duration<double, std::milli>(clock::now() - t0_).count(); }
  double elapsed_us() const { return std::chrono::duration<double, std::micro>(clock::now() - t0_).count(); }

private:
  using clock = std::chrono::steady_clock;
  clock::time_point t0_;
};

struct ScopeTimer {
  explicit ScopeTimer(std::string n) : name(std::move(n)) {}
  ~ScopeTimer() { std::cerr << "[TIMER] " << name << ": " << static_cast<long long>(t.elapsed_ms()) << " ms\n"; }
  std::string name;
  Timer t;
};

}  // namespace cs
