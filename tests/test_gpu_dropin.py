"""GPU: the reference's own callers (tools/*.cpp, tests/*.cpp), compiled UNMODIFIED against the
drop-in header + libcs_b200.so (tools/build_ref_callers.py), must print what they print when
built against the real reference — including the two assertions the reference itself trips
(tests/simple_tests.cpp:11 expects sorted locate output; tests/fm_search_tests.cpp:191 has '$'
twice), SURVEY §0."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "ref_callers", "bin")
GOLDEN = os.path.join(ROOT, "tests", "golden")
EXPECTED = json.load(open(os.path.join(GOLDEN, "ref_callers_expected.json")))["runs"]
FM = {c["name"]: c for c in json.load(open(os.path.join(GOLDEN, "golden_fm.json")))["cases"]}


def _stable(exe, text):
    keep = []
    for line in text.splitlines():
        if exe == "benchmark":
            if re.search(r"Total matches|Text size|Queries:|Pattern len", line):
                keep.append(line.strip())
        else:
            keep.append(line.rstrip())
    return keep


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("inputs")
    paths = {}
    for key, case in (("sample", "sample_txt"), ("example", "example_txt")):
        p = d / f"{key}.txt"
        p.write_bytes(bytes.fromhex(FM[case]["text_hex"]))
        paths[key] = str(p)
    return paths


@pytest.mark.skipif(not os.path.isdir(BIN), reason="build/ref_callers/bin not built (tools/build_ref_callers.py)")
@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_reference_caller_output(key, inputs):
    exp = EXPECTED[key]
    exe = os.path.join(BIN, exp["exe"])
    # stdbuf: two callers abort() on an assert; unbuffered stdout keeps what they printed before
    # (the reference build only shows it because its [TIMER] lines on cerr flush the tied cout)
    argv = ["stdbuf", "-o0", exe] + [a.format(**inputs) for a in exp["args"]]
    r = subprocess.run(argv, capture_output=True, text=True, errors="replace", timeout=600)
    assert (r.returncode == 0) == exp["returncode_is_zero"], r.stderr[-500:]
    assert _stable(exp["exe"], r.stdout) == exp["stdout"]
    err = [re.sub(r"^.*?(Assertion)", r"\1", l) for l in r.stderr.splitlines() if not l.startswith("[TIMER]")]
    assert err == exp["stderr"]


def test_cpp_batch_api(tmp_path):
    """count_batch / locate_batch of the C++ class against the single-query methods."""
    src = tmp_path / "batch.cpp"
    pkg = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200")
    src.write_text(r'''
#include "api/fm_index.hpp"
#include "csfm.h"
#include <cassert>
#include <iostream>
#include <thread>
int main(int argc, char** argv) {
  (void)argc;
  std::string text;
  for (int i = 0; i < 5000; ++i) text += "the quick brown fox jumps over the lazy dog "[(i * 7 + i / 3) % 44];
  text += '$';
  cs::BuildParams p; p.ssa_stride = 8;
  cs::FMIndex idx = cs::FMIndex::build_from_text(text, p);
  cs::FMIndex copy = idx;  // copies share the device index
  std::vector<std::string> pats = {"the", "q", "", "zzz", "o", "fox jumps", text.substr(100, 30)};
  auto counts = copy.count_batch(pats);
  std::vector<std::string_view> views(pats.begin(), pats.end());
  auto loc = idx.locate_batch(views, 17);
  for (size_t i = 0; i < pats.size(); ++i) {
    assert(counts[i] == idx.count(pats[i]));
    auto one = idx.locate(pats[i], 17);
    assert(loc.status[i] == 0);
    assert(loc.offsets[i + 1] - loc.offsets[i] == one.size());
    for (size_t k = 0; k < one.size(); ++k) assert(loc.positions[loc.offsets[i] + k] == one[k]);
    for (auto pos : one) assert(text.compare(pos, pats[i].size(), pats[i]) == 0);
  }
  assert(idx.extract(4, 5) == text.substr(4, 5) && idx.extract(text.size() + 1, 3).empty());
  bool threw = false;
  try { cs::FMIndex::open_directory("x"); } catch (const std::runtime_error&) { threw = true; }
  assert(threw);
  cs::FMIndex bad = cs::FMIndex::build_from_text("aaaa", cs::BuildParams{});
  threw = false;
  try { bad.locate("a"); } catch (const std::runtime_error& e) { threw = std::string(e.what()) == "locate: LF walk exceeded text length"; }
  assert(threw);
  // .csidx round trip: save -> load (one read + one host->device copy) -> same answers
  const std::string path = std::string(argv[1]) + "/index.csidx";
  idx.save(path);
  cs::FMIndex back = cs::FMIndex::load(path);
  assert(back.size() == idx.size());
  auto counts2 = back.count_batch(pats);
  assert(counts2 == counts);
  auto loc2 = back.locate_batch(views, 17);
  assert(loc2.offsets == loc.offsets && loc2.positions == loc.positions && loc2.status == loc.status);
  assert(back.extract(4, 5) == text.substr(4, 5));
  // a file without its TEXT section: extract() reads the text back out of the device index (fm_index.cpp:163-167)
  const std::string path2 = std::string(argv[1]) + "/index_notext.csidx";
  idx.save(path2, false);
  cs::FMIndex lean = cs::FMIndex::load(path2);
  assert(lean.count_batch(pats) == counts);
  assert(lean.extract(0, text.size() + 10) == text && lean.extract(4, 5) == text.substr(4, 5));
  assert(lean.extract(text.size() - 3, 100) == text.substr(text.size() - 3) && lean.extract(text.size(), 3).empty());
  // const query methods from several threads at once (each thread other than the builder gets its own alias handle
  // over the same device blob: own streams and workspaces, no copy), like the reference's stateless readers
  {
    std::vector<std::thread> workers;
    std::vector<int> ok(6, 0);
    for (int t = 0; t < 6; ++t)
      workers.emplace_back([&, t] {
        bool good = true;
        for (int rep = 0; rep < 20; ++rep) {
          good = good && idx.count_batch(pats) == counts;
          good = good && idx.count(pats[(t + rep) % pats.size()]) == counts[(t + rep) % pats.size()];
          auto l3 = idx.locate_batch(views, 17);
          good = good && l3.positions == loc.positions && l3.offsets == loc.offsets;
        }
        ok[t] = good;
      });
    for (auto& w : workers) w.join();
    for (int t = 0; t < 6; ++t) assert(ok[t]);
  }
  // replicas in one process: one per visible GPU (at least two handles, on the same device if there is only one)
  int ndev = 0;
  assert(csfm_device_count(&ndev) == 0 && ndev >= 1);
  std::vector<cs::FMIndex> reps = {idx};
  for (int d = 1; d < (ndev > 1 ? ndev : 2); ++d) reps.push_back(idx.replicate_to(d % ndev));
  assert(reps[1].device() == 1 % ndev && reps[1].handle() != idx.handle());
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> offs = {0};
  std::vector<std::string> many;
  for (int i = 0; i < 3001; ++i) many.push_back(text.substr((i * 37) % 4000, 1 + i % 9));
  for (auto& s : many) { bytes.insert(bytes.end(), s.begin(), s.end()); offs.push_back(bytes.size()); }
  std::vector<uint64_t> sharded(many.size()), single(many.size());
  cs::FMIndex::count_batch_sharded(reps, bytes.data(), offs.data(), many.size(), sharded.data());
  idx.count_batch(bytes.data(), offs.data(), many.size(), single.data());
  assert(sharded == single);
  cs::FMIndex big = cs::FMIndex::build_from_text(text, p, CSFM_BUILD_LARGE_TABLE);
  assert(big.count_batch(pats) == counts);
  std::cout << "ok " << counts[0] << "\n";
  return 0;
}
''')
    exe = tmp_path / "batch"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++20", "-O1", "-I" + os.path.join(pkg, "host", "src"), "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                    "-L" + os.path.join(pkg, "host"), "-lcs_b200", "-L" + pkg, "-lcsfm",
                    "-Wl,-rpath," + os.path.join(pkg, "host"), "-Wl,-rpath," + pkg], check=True)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    assert r.stdout.startswith("ok ")
