#!/usr/bin/env python
"""bench.py — headline benchmark of the FM-index count() hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5] [--large-table] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default c3 = BASELINE.json configs[2], the configuration the metric's target is quoted
on): 2^30-byte synthetic text over bytes 1..255 + 0x00 terminator (sigma = 256, 8 wavelet levels in
the reference), text-sampled patterns of length 8..32, processed in 1 M-pattern batches. A "step" is
one pass of count_batch over one 1 M-pattern batch per GPU (weak scaling: every rank runs its own
batches against its own replica of the index; the index is built once on rank 0 and replicated with
one NCCL broadcast of the device blob).

One JSON line on stdout (rank 0):
  value      whole-job count queries/s with the batches already resident in HBM (CUDA events)
  e2e        the same through the streaming host-pointer C ABI (csfm_count_batch_submit_len8/_wait):
             pinned host buffers, H2D of lengths + patterns and D2H of the counts inside the timed
             region, three steps in flight; the u32-offset, u64-offset and synchronous forms beside it
  roofline   executed algorithmic bytes (rank steps x 2 x L x line, half steps, table lookups, text
             verifications) over the mean kernel duration, against the measured HBM copy
             peak (MEASURED_PEAKS.json); `traffic` from the committed ncu capture; the stepping-only
             kernel and the large-table index on the same batches as sub-objects
  cpu_baseline  the UNMODIFIED reference cs::FMIndex::count (oracle/_ref/libcsref.so) on all host
             cores over a bounded sample of the same batch, checked bit-exact against the GPU
  construction  build of the main index: seconds, suffixes/s, sorting rounds, radix passes
  locate     configs[3]: occurrences/s of locate_batch on every rank (weak), with its own roofline,
             host-pointer e2e, resident-SA variant, walk-length histogram and reference baseline

--impl reference times only that CPU reference (rank 0; other ranks exit 0).
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: text kind/seed, pattern lengths, seeds
    "c3": dict(n_log2=30, kind="byte", seed_text=3, len_lo=8, len_hi=32, seed_len=4, seed_pos=5, batch=1_000_000,
               stride=32, desc="C3: 2^30 B text sigma=256, text-sampled patterns len 8..32, 1M-pattern batches"),
    "c2": dict(n_log2=26, kind="dna", seed_text=1, len_lo=20, len_hi=20, seed_len=0, seed_pos=2, batch=1_000_000,
               stride=32, desc="C2: 2^26 B DNA+$ text, text-sampled patterns len 20, 1M-pattern batches"),
    # configs[4]: the index is BUILT on rank 0 (suffix array, BWT, levels: `construction`) and broadcast; the
    # count sweep is steps x ranks 1M-pattern batches (100 steps on 1 GPU = the 100M patterns of the config)
    "c5": dict(n=4_000_000_000, n_log2=32, kind="dna", seed_text=7, len_lo=20, len_hi=20, seed_len=0, seed_pos=11, batch=1_000_000,
               stride=32, desc="C5: 4e9 B DNA+$ text built on the GPU, text-sampled patterns len 20, 1M-pattern batches"),
}


def text_size(args, wl):
    return (1 << args.n_log2) if args.n_log2 else wl.get("n", 1 << wl["n_log2"])
METRIC = "count queries/sec"
NB = 8  # distinct resident batches cycled through the timed steps (8 x ~28 MB > L2 with the 1.15 GB index)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Exactly ONE line may reach stdout (the JSON line). Libraries write banners to fd 1 (NCCL prints
# its version there), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
# private duplicate of the original stdout.
_REAL_STDOUT = None


def claim_stdout():
    """Called once by main(): from here on everything written to fd 1 goes to stderr, except emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    claim_stdout()
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                phys = int(vis.split(",")[torch_device_index]) if vis and vis.split(",")[0].isdigit() else torch_device_index
                self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)  # the timed region of `value` lasts ~30 ms: sample it densely

    def __enter__(self):
        if self._h is not None:
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._h is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class _DevMem:
    """Exposes a raw device pointer to torch.as_tensor (zero-copy view of the index blob)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


L2_RESIDENT_BYTES = 126 << 20   # an index up to the L2 size is served mostly from it (profiles/r2_l2_sweep_probe.json: graceful decay past 64 MiB)
RANDOM_FETCH_CEILING = 37.3e9    # random 128-byte line fetches/s out of HBM (profiles/r1_gather_probe.json, 4 GiB buffer)


_ALL_CPUS = None


class all_host_cpus:
    """The CPU baselines use every host core: lift the NUMA-local restriction of this rank while they run."""

    def __enter__(self):
        self.saved = None
        try:
            if _ALL_CPUS:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, _ALL_CPUS)
        except Exception:
            self.saved = None

    def __exit__(self, *exc):
        try:
            if self.saved:
                os.sched_setaffinity(0, self.saved)
        except Exception:
            pass


def pin_to_gpu_numa_node(local_rank):
    """Best effort: run this rank's host thread (and so its pinned allocations, first touch) on the CPUs NVML reports as
    local to its GPU. A no-op on boxes whose GPUs all hang off one NUMA node (the B200 boxes of this pool)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < ncpu}
        global _ALL_CPUS
        allowed = os.sched_getaffinity(0)
        _ALL_CPUS = set(allowed)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        return {"cpus": len(use), "of": len(allowed), "source": "nvmlDeviceGetCpuAffinity", "restricted": len(use) < len(allowed)}
    except Exception as e:  # pragma: no cover
        return {"unavailable": repr(e)}


def kernel_sources_sha16():
    """Identifies the kernel sources a committed ncu capture belongs to (.git does not travel to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".hpp")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def committed_traffic(key, kernel_name):
    """dram__bytes_read + write per launch from profiles/count_kernel_traffic.json (written by tools/ncu_traffic.py from an
    `ncu --set full` capture) -- but only when that capture was taken from the kernel sources in this tree and names the
    kernel this run launches; a stale or foreign capture must not decorate a new kernel. -> (bytes | None, meta)"""
    tp = os.path.join(ROOT, "profiles", "count_kernel_traffic.json")
    try:
        doc = json.load(open(tp))
        ent = doc.get("captures", {}).get(key)
        if not ent:
            return None, {"valid": False, "why": f"no capture for {key}"}
        sha = kernel_sources_sha16()
        base = kernel_name.split("<")[0]
        if doc.get("source_sha16") != sha:
            return None, {"valid": False, "why": f"capture belongs to sources {doc.get('source_sha16')}, this tree is {sha}"}
        if base not in ent.get("kernel", ""):
            return None, {"valid": False, "why": f"capture is of {ent.get('kernel')}, this run launches {kernel_name}"}
        meta = {"valid": True, "source_sha16": sha}
        meta.update({k: v for k, v in ent.items() if k != "dram_bytes_per_launch"})
        return float(ent["dram_bytes_per_launch"]), meta
    except Exception as e:
        return None, {"valid": False, "why": repr(e)}


def l2_random_peak(index_bytes, line_bytes):
    """Measured random-fetch ceiling (GB/s of `line_bytes` units, dependent chains) for a working set of this size."""
    try:
        doc = json.load(open(os.path.join(ROOT, "profiles", "r2_l2_sweep_probe.json")))
        rows = [r for r in doc["results"] if r["ctas_per_sm"] == 8 and r["unit_bytes"] == line_bytes]
        rows.sort(key=lambda r: r["buffer_mib"])
        prev = None
        for r in rows:
            if (r["buffer_mib"] << 20) >= index_bytes:
                if prev is None or (r["buffer_mib"] << 20) == index_bytes:
                    return float(r["gbytes_per_s"]), (f"measured: profiles/r2_l2_sweep_probe.json, {r['buffer_mib']} MiB working set, "
                                                     f"{line_bytes}-byte units, {r['gunits_per_s']:.0f} G fetches/s")
                # between two rows of the sweep: linear in the working-set size (the rate falls steeply beyond the L2)
                t = (index_bytes - (prev["buffer_mib"] << 20)) / float((r["buffer_mib"] - prev["buffer_mib"]) << 20)
                gbs = prev["gbytes_per_s"] + t * (r["gbytes_per_s"] - prev["gbytes_per_s"])
                gun = prev["gunits_per_s"] + t * (r["gunits_per_s"] - prev["gunits_per_s"])
                return float(gbs), (f"measured: profiles/r2_l2_sweep_probe.json, interpolated between the {prev['buffer_mib']} and "
                                    f"{r['buffer_mib']} MiB rows for a {index_bytes / 2**20:.0f} MiB working set, {line_bytes}-byte units, "
                                    f"{gun:.0f} G fetches/s")
            prev = r
        if prev is not None:  # beyond the sweep: out of HBM the rate no longer depends on the size
            return float(prev["gbytes_per_s"]), (f"measured: profiles/r2_l2_sweep_probe.json, {prev['buffer_mib']} MiB working set (the largest of "
                                                f"the sweep), {line_bytes}-byte units, {prev['gunits_per_s']:.0f} G fetches/s")
    except Exception:
        pass
    return 144.0 * line_bytes, "fallback: 144 G random 128-byte fetches/s measured on a 32 MiB working set (round 1 probe)"


def count_working_set(info):
    """Bytes a count call keeps coming back to: the level lines and the k-mer (+ half-step) table — not the suffix-array
    samples, which only locate reads."""
    ws = int(info.blocks_per_level) * int(info.line_bytes) * int(info.levels)
    k = int(info.kmer_k)
    if k:
        radix = min(int(info.sigma), 4) if int(info.layout) == 3 else int(info.sigma)
        entries = radix ** k
        ws += entries * (4 if int(info.text_check) else 8) + (entries * 16 * 8 if int(info.half_table) else 0)
    return ws


def make_text(wl, n, device):
    from csfm_b200 import workloads as w
    return (w.byte_text_torch if wl["kind"] == "byte" else w.dna_text_torch)(n, wl["seed_text"], device)


def make_batch(wl, text, first, npat):
    from csfm_b200 import workloads as w
    return w.sampled_patterns_torch(text, npat, wl["len_lo"], wl["len_hi"], wl["seed_len"], wl["seed_pos"], first=first)


def planes_from_blob(blob: np.ndarray):
    """Device blob (8 levels of 64-byte lines) -> the reference's packed u64 bit planes."""
    magic, version, levels, n, sigma, stride, nsamp, nblk, off_levels, level_stride, off_ssa, total = struct.unpack_from(
        "<8sIIQIIQQQQQQ", blob, 0)
    assert magic == b"CSFMDEV1" and levels == 8
    nwords = (n + 63) // 64
    planes = []
    for l in range(8):
        lines = blob[off_levels + l * level_stride: off_levels + l * level_stride + nblk * 64].view(np.uint32).reshape(nblk, 16)
        payload = np.ascontiguousarray(lines[:, 1:]).reshape(-1)
        if payload.size % 2:
            payload = np.concatenate([payload, np.zeros(1, np.uint32)])
        planes.append(payload.view(np.uint64)[:nwords].copy())
    return n, planes


def build_reference_index(fm, text, wl, with_locate=False):
    """UNMODIFIED cs::FMIndex (oracle/_ref) over the same text. The O(n^2 log n) build_sa_naive cannot
    run at this size, so the BWT comes from the GPU builder (untimed setup) and the reference's own
    BitVector::build_from_words builds its rank directory over the bit planes."""
    import torch
    import oracle
    n = text.numel()
    idx8 = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=wl["stride"]),
                                             device=text.device.index, flags=fm.BUILD_NO_COMPACT | fm.BUILD_LAYOUT_BINARY64)
    blob = idx8.blob_to_host()
    bwt = ssa = None
    if with_locate:  # the reference's LF reads bwt_[i] and its locate reads ssa_ (fm_index.hpp:62-66, fm_index.cpp:141-152)
        bwt, ssa = idx8.bwt(), idx8.ssa()
    idx8.close()
    n2, planes = planes_from_blob(blob)
    del blob
    assert n2 == n
    hist = np.zeros(256, np.uint64)  # C array from the text's own histogram, independent of the engine
    for s0 in range(0, n, 1 << 27):
        hist += torch.bincount(text[s0: s0 + (1 << 27)].int(), minlength=256).cpu().numpy().astype(np.uint64)
    Carr = np.zeros(257, np.uint32)
    Carr[1:] = np.cumsum(hist).astype(np.uint32)
    return oracle.RefIndex(planes=planes, C_array=Carr, n=n, stride=wl["stride"], bwt=bwt, ssa=ssa)


def reference_timed_sample(ref, data, offs, budget_s, nthreads, min_queries=None):
    """Runs the reference's count() over the first q queries of a batch, q sized to ~budget_s."""
    q0 = min(offs.size - 1, max(nthreads, min_queries or 0, 8))
    t = time.perf_counter()
    out0 = ref.count_batch(data, offs[: q0 + 1], nthreads=nthreads)
    dt0 = time.perf_counter() - t
    rate = q0 / max(dt0, 1e-9)
    q = int(min(offs.size - 1, max(q0, rate * budget_s)))
    if q <= q0 * 1.5:
        return out0, q0, dt0
    t = time.perf_counter()
    out = ref.count_batch(data, offs[: q + 1], nthreads=nthreads)
    return out, q, time.perf_counter() - t


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    import csfm_b200 as fm
    import oracle
    wl = dict(WORKLOADS[args.workload])
    n = text_size(args, wl)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    text = make_text(wl, n, dev)
    ref = build_reference_index(fm, text, wl)
    nthreads = os.cpu_count() or 1
    bytes_d, offs_d = make_batch(wl, text, 0, min(wl["batch"], 200_000))
    data, offs = bytes_d.cpu().numpy(), offs_d.cpu().numpy().astype(np.uint64)
    # Size the per-step sample so that warmup + steps end in ~2 minutes whatever K is. Each query is
    # one std::thread's work, so a step with fewer queries than cores keeps only that many cores busy
    # (reported in `cores`); with few steps every core works.
    _, q0, dt0 = reference_timed_sample(ref, data, offs, 0.0, nthreads)
    rate = q0 / dt0  # queries/s with all cores busy
    per_step = int(rate * 120.0 / (args.steps + args.warmup))
    per_step = max(1, min(per_step, (offs.size - 1) // 2))
    threads_used = min(nthreads, per_step)
    total_q, total_t, cursor = 0, 0.0, 0
    for i in range(args.warmup + args.steps):
        lo = cursor % (offs.size - 1 - per_step)
        o = offs[lo: lo + per_step + 1]
        t = time.perf_counter()
        ref.count_batch(data, o, nthreads=threads_used)
        dt = time.perf_counter() - t
        cursor += per_step
        if i >= args.warmup:
            total_q += per_step
            total_t += dt
    qps = total_q / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": wl["desc"], "n": n, "queries_per_step": per_step,
                   "note": "reference cs::FMIndex::count verbatim on host cores; its index is injected from the GPU-built "
                           "BWT (untimed setup) because build_sa_naive is O(n^2 log n)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads_used, "kind": "reference",
                         "sample": f"{per_step} queries per step x {args.steps} steps of the first batch on {threads_used} of "
                                   f"{nthreads} host threads (all-core calibration: {rate:.2f} q/s)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)



# ------------------------------------------------------------------------------------------------
# locate leg (BASELINE.json configs[3]: DNA text, ssa_stride 32, frequent text-sampled patterns)
# ------------------------------------------------------------------------------------------------
def locate_cpu_baseline(fm, idx, text, bytes_d, offs_d, plen, budget_s):
    """The unmodified reference's locate() on the host cores over a bounded sample (a few patterns, small
    limit: every LF step of the reference rescans whole bit planes), checked against the GPU."""
    import oracle
    nthreads = os.cpu_count() or 1
    t0 = time.perf_counter()
    ref = build_reference_index(fm, text, {"stride": 32}, with_locate=True)
    t_inject = time.perf_counter() - t0
    q, lim = nthreads, 2
    data = bytes_d[: q * plen].cpu().numpy()
    offs = offs_d[: q + 1].cpu().numpy().astype(np.uint64)
    t0 = time.perf_counter()
    tot, out_n, pos = ref.locate_batch(data, offs, limit=lim, nthreads=nthreads, keep_positions=True)
    dt = time.perf_counter() - t0
    rate = tot / max(dt, 1e-9)
    lim2 = int(min(64, max(lim, rate * budget_s / q)))
    if lim2 >= 2 * lim:
        lim = lim2
        t0 = time.perf_counter()
        tot, out_n, pos = ref.locate_batch(data, offs, limit=lim, nthreads=nthreads, keep_positions=True)
        dt = time.perf_counter() - t0
    g_offs, g_pos, g_status = idx.locate_batch(data, offs, limit=lim)
    ok = bool((g_status == 0).all()) and int(g_offs[-1]) == int(tot)
    for k in range(q):
        ok = ok and bool((g_pos[int(g_offs[k]): int(g_offs[k + 1])] == pos[k * lim: k * lim + int(out_n[k])]).all())
    return {"value": tot / dt, "unit": "occurrences/s", "cores": nthreads, "kind": "reference",
            "sample": f"{q} patterns, limit {lim}: {tot} occurrences in {dt:.1f} s on {nthreads} std::threads "
                      f"(reference index injected in {t_inject:.0f} s)",
            "bit_exact_vs_gpu": ok}


def measure_locate(fm, dev, rank=0, world=1, n_log2=28, npat=1_000_000, plen=10, limit=100_000, iters=5, cpu_budget=0.0):
    """locate_batch on device-resident inputs/outputs: occurrences/s, checked by re-reading the text.

    Weak scaling like the count leg: rank 0 builds the index, one broadcast replicates it, every rank
    locates its own `npat` patterns; value = occurrences of all ranks / max over ranks of the device time.
    Beside it (rank-local, reported by rank 0): the same batch through the host-pointer call
    csfm_locate_batch (`e2e`) and on an index that keeps its whole suffix array (`resident_sa`)."""
    import torch
    import torch.distributed as dist
    from csfm_b200 import workloads as w
    n = 1 << n_log2
    text = w.dna_text_torch(n, 6, dev)
    build_s = bcast_ms = None
    idx = None
    if rank == 0:
        t0 = time.perf_counter()
        idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=dev.index)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
        idx, bcast_ms = fm.parallel.replicate_index(idx, dev, src=0)
    info = idx.info()
    bytes_d, offs_d = w.sampled_patterns_torch(text, npat, plen, plen, 0, 8, first=rank * npat)
    stream = torch.cuda.Stream(device=dev)
    d_offs_out = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(npat, dtype=torch.int32, device=dev)
    total = idx.locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, d_offs_out.data_ptr(), 0, 0,
                                    d_status.data_ptr(), stream.cuda_stream)
    d_pos = torch.zeros(max(1, total), dtype=torch.int64, device=dev)

    def run(ix=None, pos=None):
        return (ix or idx).locate_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), npat, limit, d_offs_out.data_ptr(),
                                               (d_pos if pos is None else pos).data_ptr(), total, d_status.data_ptr(),
                                               stream.cuda_stream)

    def timed(fn, k, all_ranks=True):
        for _ in range(2):
            fn()
        stream.synchronize()
        if world > 1 and all_ranks:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                fn()
            e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / k

    idx.set_instrumentation(1)
    run()
    stream.synchronize()
    lf_steps = int(idx.last_call_stats().lf_steps)
    idx.set_instrumentation(0)
    # correctness: every reported position really starts an occurrence of its pattern; no failed query
    assert int(d_status.max().item()) == 0
    sample = torch.randint(0, total, (min(total, 2_000_000),), device=dev)
    q_of = torch.searchsorted(d_offs_out, sample, right=True) - 1
    ok = torch.ones(sample.numel(), dtype=torch.bool, device=dev)
    for k in range(plen):
        ok &= text[d_pos[sample] + k] == bytes_d[offs_d[q_of] + k]
    positions_ok = bool(ok.all().item())
    counts_ok = bool(((d_offs_out[1:] - d_offs_out[:-1]) >= 1).all().item())
    ms = timed(run, iters)
    ms_local, total_all, lf_all = ms, total, lf_steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([total, lf_steps, int(positions_ok and counts_ok)], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ms, total_all, lf_all = float(t[0]), int(c[0]), int(c[1])
        positions_ok = counts_ok = int(c[2]) == world
    if rank != 0:
        idx.close()
        return None
    L, lb = int(info.levels), int(info.line_bytes)
    # executed fetches are exact here: every LF step loads exactly L level lines, every occurrence one sample sector
    # and writes 8 bytes (sum over occurrences of w x L x line + 32 + 8). With position samples the walk also
    # reads the line of the marked row it stops at (that is where the mark bit lives): w + 1 lines per occurrence.
    line_loads = (lf_all + (total_all if int(info.position_samples) else 0)) * L
    alg = line_loads * lb + total_all * 40
    peak, peak_src = measured_peak()
    t_s = ms / 1e3
    kname = {1: "walk_kernel", 2: "walk2_kernel", 3: "walk3_kernel"}.get(int(info.layout), "walk_kernel")
    traffic, tmeta = committed_traffic("c4_walk", kname)
    if traffic and tmeta.get("units_per_launch"):
        # the capture is of a smaller batch: DRAM bytes per occurrence x this batch's occurrences (per GPU)
        traffic = traffic / float(tmeta["units_per_launch"]) * (total_all / world)
        tmeta["scaled_to_occurrences"] = total_all / world
    elif traffic:
        traffic, tmeta = None, {"valid": False, "why": "capture does not say how many occurrences it walked"}
    # what the walk touches: the level lines and the samples (the k-mer table is only read by the count pass)
    walk_set = int(info.blocks_per_level) * lb * L + int(info.nsamp) * 4
    in_l2 = walk_set <= L2_RESIDENT_BYTES and int(info.blob_bytes) <= L2_RESIDENT_BYTES
    ws_peak, ws_src = l2_random_peak(walk_set, lb)
    achieved_exec = alg / world / t_s / 1e9
    if in_l2:
        roof = {"bound": "l2", "achieved": achieved_exec, "peak": ws_peak, "frac": achieved_exec / ws_peak, "peak_source": ws_src,
                "frac_basis": "executed line fetches x line bytes over the measured random-fetch ceiling for a working set of this size"}
    else:
        ach = (traffic / t_s / 1e9) if traffic else achieved_exec
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "frac": ach / peak, "peak_source": peak_src,
                "frac_basis": "dram traffic (ncu) / duration" if traffic else
                "executed line fetches x line bytes / duration (part of them hit the L2: no valid ncu capture to subtract them)",
                "random_fetch_ceiling_at_this_working_set_gbs": ws_peak, "frac_of_that_ceiling": achieved_exec / ws_peak,
                "ceiling_source": ws_src}
    roof.update({"kernel": kname, "unit": "GB/s", "traffic": traffic, "traffic_capture": tmeta,
                 "executed": {"bytes_per_batch": alg / world, "gbs": achieved_exec, "level_lines_per_s": line_loads / world / t_s},
                 "note": "per GPU; time covers count pass + scan + expand + walk"})
    out = {"metric": "locate occurrences/sec", "value": total_all / t_s, "unit": "occurrences/s", "n_gpus": world,
           "ms_per_batch": ms, "scaling": "weak",
           "config": {"workload": f"C4: 2^{n_log2} B DNA+$ text, ssa_stride 32, {npat} text-sampled patterns len {plen} per GPU, limit {limit}",
                      "levels": L, "line_bytes": lb, "layout": int(info.layout), "position_samples": int(info.position_samples),
                      "index_bytes": int(info.blob_bytes),
                      "walk_working_set_bytes": walk_set, "kmer_k": int(info.kmer_k), "index_build_s": build_s,
                      "index_broadcast_ms": bcast_ms},
           "occurrences_per_batch": int(total_all), "lf_steps_per_occurrence": lf_all / max(1, total_all),
           "roofline": roof,
           "checks": {"positions_verified_against_text": positions_ok, "all_counts_ge_1": counts_ok, "failed_queries": 0}}

    # ---- the same batch through the host-pointer call (pinned buffers; H2D and D2H inside the timed region)
    try:
        h_bytes = bytes_d.cpu().pin_memory()
        h_offs = offs_d.cpu().pin_memory()
        h_out_offs = torch.zeros(npat + 1, dtype=torch.int64).pin_memory()
        h_pos = torch.zeros(total, dtype=torch.int64).pin_memory()
        h_status = torch.zeros(npat, dtype=torch.int32).pin_memory()
        import ctypes
        tot_c = ctypes.c_uint64()
        L_ = fm.lib()

        def host_call():
            rc = L_.csfm_locate_batch(idx._h, h_bytes.data_ptr(), h_offs.data_ptr(), npat, limit, h_out_offs.data_ptr(),
                                      h_pos.data_ptr(), total, h_status.data_ptr(), ctypes.byref(tot_c))
            if rc != 0:
                raise RuntimeError(L_.csfm_last_error().decode())

        host_call()
        host_call()
        t0 = time.perf_counter()
        for _ in range(3):
            host_call()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / 3
        out["e2e"] = {"value": total / (e2e_ms / 1e3), "unit": "occurrences/s", "ms_per_batch": e2e_ms, "n_gpus": 1,
                      "h2d_bytes_per_step": int(h_bytes.numel() + 8 * h_offs.numel()),
                      "d2h_bytes_per_step": int(8 * total + 8 * (npat + 1) + 4 * npat),
                      "api": "csfm_locate_batch (host pointers, pinned; positions are 8 B each, so the D2H copy of the "
                             "result is the larger part of the step)",
                      "equals_device_path": bool((h_pos.numpy() == d_pos.cpu().numpy()).all())}
        del h_pos
    except Exception as e:  # pragma: no cover
        out["e2e"] = {"unavailable": repr(e)}

    # ---- an index that keeps its whole suffix array (4 n bytes more): a row's position is one gather
    try:
        idx_sa = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=32), device=dev.index,
                                                   flags=fm.BUILD_FORCE_TEXT_CHECK | fm.BUILD_KEEP_SA)
        d_pos2 = torch.zeros_like(d_pos)
        ms_sa = timed(lambda: run(idx_sa, d_pos2), iters, all_ranks=False)  # rank 0 only
        out["resident_sa"] = {"value": total / (ms_sa / 1e3), "unit": "occurrences/s", "ms_per_batch": ms_sa, "n_gpus": 1,
                              "index_bytes": int(idx_sa.info().blob_bytes),
                              "positions_equal_walk": bool(torch.equal(d_pos, d_pos2)),
                              "note": "CSFM_BUILD_FORCE_TEXT_CHECK: text + full suffix array ride in the index blob, locate "
                                      "reads SA[row] instead of walking LF to a sampled row (same positions, same order)"}
        # walk-length histogram of the batch, from the resident suffix array: the walk of an
        # occurrence at text position p ends at the nearest sampled position at or before p (cyclically), so
        # its length is the distance to it. The sum must equal the LF steps the walk kernel counted.
        try:
            if int(info.position_samples):
                # marked form: the samples sit at the text positions that are multiples of the stride
                wl = d_pos % 32
            else:
                sa = torch.as_tensor(_DevMem(idx_sa.sa_device_ptr(), 4 * n), device=dev).view(torch.int32)  # n < 2^31 here
                sampled = sa[::32].to(torch.int64)                 # SA[k * stride]: the positions whose rows are sampled
                mark = torch.full((n,), -1, dtype=torch.int64, device=dev)
                mark[sampled] = sampled
                prev = torch.cummax(mark, 0).values               # nearest sampled position <= q, -1 if none
                del mark
                pp = prev[d_pos]
                wl = torch.where(pp >= 0, d_pos - pp, d_pos + (n - int(sampled.max())))
                del prev, pp
            edges = [0, 1, 8, 16, 32, 64, 128, 256]
            hist = torch.bincount(torch.bucketize(wl, torch.tensor(edges[1:], device=dev), right=True), minlength=len(edges))
            out["walk_lengths"] = {"mean": float(wl.double().mean()), "max": int(wl.max()),
                                   "p50": int(wl.kthvalue(max(1, wl.numel() // 2)).values),
                                   "p99": int(wl.kthvalue(max(1, int(wl.numel() * 0.99))).values),
                                   "bucket_edges": edges, "bucket_counts": [int(x) for x in hist.cpu()],
                                   "sum_equals_lf_steps_counted": int(wl.sum()) == lf_steps,
                                   "note": "LF steps per occurrence of this rank's batch; buckets are [edge_i, edge_i+1), the last one open"}
            del wl
        except Exception as e:  # pragma: no cover
            out["walk_lengths"] = {"unavailable": repr(e)}
        idx_sa.close()
        del d_pos2
    except Exception as e:  # pragma: no cover
        out["resident_sa"] = {"unavailable": repr(e)}

    if cpu_budget > 0:
        try:
            with all_host_cpus():
                out["cpu_baseline"] = locate_cpu_baseline(fm, idx, text, bytes_d, offs_d, plen, cpu_budget)
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "unit": "occurrences/s", "kind": "reference", "sample": f"unavailable: {e!r}"}
    idx.close()
    return out

# ------------------------------------------------------------------------------------------------
# engine arm
# ------------------------------------------------------------------------------------------------
def count_leg(args, wl_name, rank, world, local_rank, steps, warmup, full):
    """One count workload end to end: text, index (built on rank 0, one broadcast), resident batches, the
    device-resident timed region, the host-pointer e2e region, roofline, CPU baseline. `full` adds the
    comparison legs of the headline line (stepping-only kernel, large-table index, every e2e form).
    Returns the JSON object on rank 0, None elsewhere. Frees everything it allocated."""
    import torch
    import torch.distributed as dist
    import csfm_b200 as fm

    wl = dict(WORKLOADS[wl_name])
    n = text_size(args, wl) if wl_name == args.workload else wl.get("n", 1 << wl["n_log2"])
    batch = args.batch or wl["batch"]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    fm.lib()  # fail loudly if the CUDA extension is missing

    class _A:  # the leg's own step counts under the names the body uses
        pass
    largs = _A()
    largs.__dict__.update(vars(args))
    largs.steps, largs.warmup, largs.workload = steps, warmup, wl_name
    args = largs

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- text + index (rank 0 builds, one broadcast replicates) ------------------------------------
    t0 = time.perf_counter()
    text = make_text(wl, n, dev)
    torch.cuda.synchronize()
    t_text = time.perf_counter() - t0
    build_s = bcast_ms = None
    if rank == 0:
        torch.cuda.empty_cache()  # the builder allocates with cudaMalloc: give it what torch's cache is holding
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=wl["stride"]), device=local_rank,
                                                flags=(fm.BUILD_LAYOUT_BINARY64 if args.layout == 1 else 0) |
                                                (fm.BUILD_LARGE_TABLE if args.large_table else 0))
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
    if world > 1:
        # warm NCCL up on a small tensor so that bcast_ms measures the transfer, not communicator setup
        dist.all_reduce(torch.zeros(1, device=dev))
        barrier()
        idx, bcast_ms = fm.parallel.replicate_index(idx if rank == 0 else None, dev, src=0)  # the one ncclBroadcast
    info = idx.info()
    L = int(info.levels)
    construction = None
    if rank == 0:
        construction = {"n": n, "seconds": build_s, "suffixes_per_s": n / build_s, "sa_rounds": int(info.sa_rounds),
                        "sa_radix_passes": int(info.sa_radix_passes),
                        "sa_pair_passes": int(info.sa_pair_passes),
                        "sort_bytes_moved_model": int(info.sa_pair_passes) * 2 * 12,
                        "note": "csfm_build_from_text_device on device-resident text: suffix array (packed-key radix sort + prefix "
                                "doubling over the unresolved suffixes), BWT, SSA, levels, k-mer and half-step tables, text sections; the model "
                                "figure is pairs sorted x 8-bit radix passes (summed over rounds) x 2 x (8 B key + 4 B suffix)"}
    log(f"[rank {rank}] n={n} levels={L} sigma={info.sigma} blob={info.blob_bytes/1e6:.1f} MB text_gen={t_text:.2f}s "
        f"build={build_s if build_s is None else round(build_s, 2)}s bcast_ms={bcast_ms}")

    # ---- batches: NB distinct batches per rank, resident on the device and mirrored in pinned host memory
    d_batches, h_batches = [], []
    for b in range(NB):
        first = (b * world + rank) * batch
        bytes_d, offs_d = make_batch(wl, text, first, batch)
        d_batches.append((bytes_d, offs_d))
        hb = torch.empty(bytes_d.numel(), dtype=torch.uint8).pin_memory()
        ho = torch.empty(offs_d.numel(), dtype=torch.int64).pin_memory()
        hb.copy_(bytes_d)
        ho.copy_(offs_d)
        h_batches.append((hb, ho))
    d_counts = torch.zeros(batch, dtype=torch.int64, device=dev)
    h_counts = torch.zeros(batch, dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)

    def step_device(b):
        bytes_d, offs_d = d_batches[b % NB]
        idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), batch, d_counts.data_ptr(), 0, stream.cuda_stream)

    # exact executed-step counts per batch (S of the roofline model) + correctness properties
    idx.set_instrumentation(1)
    steps_per_batch, lookups_per_batch, checks_per_batch, halves_per_batch, lines_per_batch = [], [], [], [], []
    for b in range(NB):
        step_device(b)
        stream.synchronize()
        steps_per_batch.append(int(idx.last_call_stats().search_steps))
        lines_per_batch.append(int(idx.last_call_stats().line_fetches))
        kernels_per_step = int(idx.last_call_stats().kernel_launches)
        lookups_per_batch.append(int(idx.last_call_stats().table_lookups))
        checks_per_batch.append(int(idx.last_call_stats().text_checks))
        halves_per_batch.append(int(idx.last_call_stats().half_steps))
        if b == 0:
            c0 = d_counts.cpu().numpy().copy()
            assert (c0 >= 1).all(), "text-sampled patterns must occur at least once"
    idx.set_instrumentation(0)

    # ---- timed region 1: device-resident inputs -----------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    stream.synchronize()
    torch.cuda.synchronize()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        with torch.cuda.stream(stream):
            for i in range(args.steps):
                evs[i].record(stream)
                step_device(i)
            evs[args.steps].record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
    barrier()
    total_ms = evs[0].elapsed_time(evs[args.steps])
    step_ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)])
    launches = args.steps * kernels_per_step  # count kernels per step (two in the two-pass form; the cursor memset is not a kernel)

    # ---- the stepping-only kernel on the same index (count2_kernel<false>): asking for the [sp,ep)
    # intervals rules the text verification out, so every query runs the plain backward search
    stepping = None
    if full and int(info.text_check) and int(info.layout) == 2:
        d_spep = torch.zeros(2 * batch, dtype=torch.int64, device=dev)

        def step_plain(b):
            bytes_d, offs_d = d_batches[b % NB]
            idx.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), batch, d_counts.data_ptr(), d_spep.data_ptr(),
                                   stream.cuda_stream)

        idx.set_instrumentation(1)
        step_plain(0)
        stream.synchronize()
        plain_steps = int(idx.last_call_stats().search_steps)
        plain_lookups = int(idx.last_call_stats().table_lookups)
        plain_lines = int(idx.last_call_stats().line_fetches)
        idx.set_instrumentation(0)
        plain_equal = bool((d_counts.cpu().numpy() == c0).all())
        K2 = min(args.steps, 40)
        for i in range(3):
            step_plain(i)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K2):
            step_plain(i)
        e1.record(stream)
        stream.synchronize()
        plain_ms = e0.elapsed_time(e1) / K2
        stepping = {"kernel": "count2_kernel<false,false>", "ms_per_launch": plain_ms, "launches_timed": K2,
                    "queries_per_s": batch / (plain_ms / 1e3), "search_steps_per_launch": plain_steps,
                    "table_lookups_per_launch": plain_lookups, "level_lines_fetched_per_launch": plain_lines,
                    "counts_equal_default_kernel": plain_equal}
        launches += K2 + 4
        del d_spep

    # ---- the same batches on an index built with a LARGE k-mer table budget (opt-in, CSFM_BUILD_LARGE_TABLE):
    # for a byte alphabet k goes from 3 to 4 (2^32 entries, 17 GB), the lookup leaves ~1 row and the query
    # goes straight to the text verification: one table line + one suffix-array line + one text line
    large = None
    if full and world == 1 and not args.no_large_table and not args.large_table and int(info.layout) == 2 and torch.cuda.mem_get_info()[0] > 100e9:
        try:
            t0 = time.perf_counter()
            idx_big = fm.FMIndex.build_from_text_device(text.data_ptr(), n, fm.BuildParams(ssa_stride=wl["stride"]), device=local_rank,
                                                        flags=fm.BUILD_LARGE_TABLE)
            torch.cuda.synchronize()
            big_build_s = time.perf_counter() - t0
            binfo = idx_big.info()
            d_counts_big = torch.zeros(batch, dtype=torch.int64, device=dev)

            def step_big(b):
                bytes_d, offs_d = d_batches[b % NB]
                idx_big.count_batch_device(bytes_d.data_ptr(), offs_d.data_ptr(), batch, d_counts_big.data_ptr(), 0, stream.cuda_stream)

            idx_big.set_instrumentation(1)
            step_big(0)
            stream.synchronize()
            bst = idx_big.last_call_stats()
            big_steps, big_lookups, big_checks = int(bst.search_steps), int(bst.table_lookups), int(bst.text_checks)
            big_lines = int(bst.line_fetches)
            idx_big.set_instrumentation(0)
            big_equal = bool((d_counts_big.cpu().numpy() == c0).all())
            K3 = min(args.steps, 100)
            for i in range(3):
                step_big(i)
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(K3):
                step_big(i)
            e1.record(stream)
            stream.synchronize()
            big_ms = e0.elapsed_time(e1) / K3
            big_bytes = big_steps * 2 * L * int(info.line_bytes) + big_lookups * 128 + big_checks * 256
            large = {"build_flag": "CSFM_BUILD_LARGE_TABLE", "kmer_k": int(binfo.kmer_k), "index_bytes": int(binfo.blob_bytes),
                     "index_build_s": big_build_s, "ms_per_launch": big_ms, "launches_timed": K3,
                     "queries_per_s": batch / (big_ms / 1e3), "search_steps_per_launch": big_steps,
                     "table_lookups_per_launch": big_lookups, "text_checks_per_launch": big_checks,
                     "level_lines_fetched_per_launch": big_lines,
                     "algorithmic_bytes_per_launch": big_bytes, "achieved_gbs": big_bytes / (big_ms / 1e3) / 1e9,
                     "counts_equal_default_index": big_equal,
                     "note": "opt-in build: the k-mer table takes up to 40 GiB of the 180 GB (k = 4 on bytes, 13 on DNA+$), a query is "
                             "three dependent fetches (table, suffix array, text); fewer executed bytes per query, so a lower "
                             "bandwidth figure at a higher q/s"}
            launches += K3 + 4
            idx_big.close()
            del d_counts_big
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            large = {"unavailable": repr(e)}

    # ---- timed region 2: end to end through the host-pointer C ABI ------------------------------------
    # The streaming form of the public API (csfm_count_batch_submit / _wait): every step copies its
    # own patterns + offsets host->device from pinned memory and its own counts device->host, all
    # inside the timed region; up to DEPTH steps are in flight so that the PCIe copies of one step
    # overlap the kernel of another.
    DEPTH = 3
    h_out = [torch.zeros(batch, dtype=torch.int64).pin_memory() for _ in range(DEPTH)]
    # compact form (csfm_count_batch_submit32): u32 offsets in, u32 counts out
    h_offs32 = []
    for hb, ho in h_batches:
        assert int(ho[-1]) < 2 ** 32
        o32 = torch.empty(ho.numel(), dtype=torch.int32).pin_memory()
        o32.copy_(ho.to(torch.int32))  # values < 2^31 here; the C side reads them as unsigned
        h_offs32.append(o32)
    h_out32 = [torch.zeros(batch, dtype=torch.int32).pin_memory() for _ in range(DEPTH)]

    # most compact form (csfm_count_batch_submit_len8): one length byte per pattern in, u32 counts out
    h_lens8 = []
    for hb, ho in h_batches:
        ln = (ho[1:] - ho[:-1])
        assert int(ln.max()) <= 255
        l8 = torch.empty(ln.numel(), dtype=torch.uint8).pin_memory()
        l8.copy_(ln.to(torch.uint8))
        h_lens8.append(l8)
    h_out8 = [torch.zeros(batch, dtype=torch.int32).pin_memory() for _ in range(DEPTH)]

    # smallest form (csfm_count_batch_submit_packed): for alphabets of at most 16 symbols the patterns travel as the
    # index's wire codes, `bits` per symbol (3 for DNA + terminator); the client holds them packed (untimed setup)
    codes_np, wire_bits = idx.pattern_codes()
    h_packed = []
    if wire_bits <= 4:
        lut = torch.from_numpy(codes_np.astype(np.int64)).to(dev)
        weights = (1 << torch.arange(8, device=dev, dtype=torch.int64))
        for bytes_d, _ in d_batches:
            sym = lut[bytes_d.long()]
            bitsm = ((sym[:, None] >> torch.arange(wire_bits, device=dev)[None, :]) & 1).reshape(-1)
            pad = (-bitsm.numel()) % 8
            if pad:
                bitsm = torch.cat([bitsm, torch.zeros(pad, dtype=bitsm.dtype, device=dev)])
            pk = (bitsm.reshape(-1, 8) * weights[None, :]).sum(1).to(torch.uint8)
            hp = torch.empty(pk.numel(), dtype=torch.uint8).pin_memory()
            hp.copy_(pk)
            h_packed.append(hp)
            del sym, bitsm, pk
    h_outp = [torch.zeros(batch, dtype=torch.int32).pin_memory() for _ in range(DEPTH)]

    def e2e_run(nsteps, form):
        tickets = []
        for i in range(nsteps):
            hb, ho = h_batches[i % NB]
            if i >= DEPTH:
                idx.count_batch_wait(tickets[i - DEPTH])  # frees output buffer i % DEPTH
            if form == "packed":
                tickets.append(idx.count_batch_submit_packed(h_packed[i % NB].data_ptr(), hb.numel(), h_lens8[i % NB].data_ptr(), batch,
                                                             h_outp[i % DEPTH].data_ptr()))
            elif form == "len8":
                tickets.append(idx.count_batch_submit_len8(hb.data_ptr(), hb.numel(), h_lens8[i % NB].data_ptr(), batch,
                                                           h_out8[i % DEPTH].data_ptr()))
            elif form == "u32":
                tickets.append(idx.count_batch_submit32(hb.data_ptr(), h_offs32[i % NB].data_ptr(), batch, h_out32[i % DEPTH].data_ptr()))
            else:
                tickets.append(idx.count_batch_submit(hb.data_ptr(), ho.data_ptr(), batch, h_out[i % DEPTH].data_ptr()))
        for t in tickets[-DEPTH:]:
            idx.count_batch_wait(t)

    e2e_steps = args.steps
    e2e_times, e2e_bytes = {}, {}
    e2e_clock_samples, e2e_clock_reasons = [], set()
    for form in (("u64", "u32", "len8") if full else ("len8",)) + (("packed",) if h_packed else ()):
        e2e_run(max(3, args.warmup), form)
        torch.cuda.synchronize()
        barrier()
        with ClockSampler(local_rank) as clocks_e2e_form:  # the e2e regions last longer: more clock samples under load
            t0 = time.perf_counter()
            e2e_run(e2e_steps, form)
            torch.cuda.synchronize()
            e2e_times[form] = time.perf_counter() - t0
        e2e_clock_samples.extend(clocks_e2e_form.samples)
        e2e_clock_reasons |= clocks_e2e_form.reasons
        barrier()
        stf = idx.last_call_stats()
        e2e_bytes[form] = (int(stf.h2d_bytes), int(stf.d2h_bytes))
    h2d32, d2h32 = e2e_bytes["len8"]
    e2e_s = e2e_times["len8"]
    e2e_s_u32, e2e_s_u64 = e2e_times.get("u32", 0.0), e2e_times.get("u64", 0.0)
    last = (e2e_steps - 1) % DEPTH
    h2d, d2h = h2d32, d2h32
    e2e_sync_ms = h2d64 = d2h64 = None
    compact_equal = None
    if full:
        compact_equal = bool((h_out32[last].numpy().astype(np.int64) == h_out[last].numpy()).all() and
                             (h_out8[last].numpy().astype(np.int64) == h_out[last].numpy()).all())
        # the synchronous call (csfm_count_batch) for comparison: one step at a time, nothing overlapped
        h_sync = torch.zeros(batch, dtype=torch.int64).pin_memory()
        L_ = fm.lib()
        for i in range(3):  # warm-up: first call creates the slice streams and sizes the workspace
            hb, ho = h_batches[i % NB]
            L_.csfm_count_batch(idx._h, hb.data_ptr(), ho.data_ptr(), batch, h_sync.data_ptr(), None)
        t0 = time.perf_counter()
        for i in range(min(20, e2e_steps)):
            hb, ho = h_batches[i % NB]
            if L_.csfm_count_batch(idx._h, hb.data_ptr(), ho.data_ptr(), batch, h_sync.data_ptr(), None) != 0:
                raise RuntimeError(L_.csfm_last_error().decode())
        e2e_sync_ms = 1e3 * (time.perf_counter() - t0) / min(20, e2e_steps)
        st = idx.last_call_stats()
        h2d64, d2h64 = int(st.h2d_bytes), int(st.d2h_bytes)
    step_device(e2e_steps - 1)  # same batch as the last end-to-end step: both paths must agree
    stream.synchronize()
    e2e_equal = bool((h_out8[last].numpy().astype(np.int64) == d_counts.cpu().numpy()).all())
    if h_packed:
        e2e_equal = e2e_equal and bool((h_outp[last].numpy().astype(np.int64) == d_counts.cpu().numpy()).all())

    # ---- the ceiling the e2e region is up against: the same bytes per step, copies only (all ranks at once) -------
    # pinned H2D of (lengths + pattern bytes) and pinned D2H of the u32 counts on two streams, no kernel
    copy_s = None
    try:
        # like the product's three async slots: step i uses stream i % 3, H2D then D2H on that stream
        h_in = [torch.zeros(h2d, dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
        d_in = [torch.empty(h2d, dtype=torch.uint8, device=dev) for _ in range(DEPTH)]
        d_o = [torch.zeros(d2h, dtype=torch.uint8, device=dev) for _ in range(DEPTH)]
        h_o = [torch.zeros(d2h, dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
        cstreams = [torch.cuda.Stream(device=dev) for _ in range(DEPTH)]

        def copies(k):
            for i in range(k):
                with torch.cuda.stream(cstreams[i % DEPTH]):
                    d_in[i % DEPTH].copy_(h_in[i % DEPTH], non_blocking=True)
                    h_o[i % DEPTH].copy_(d_o[i % DEPTH], non_blocking=True)
            for cs_ in cstreams:
                cs_.synchronize()

        copies(5)
        barrier()
        t0 = time.perf_counter()
        copies(e2e_steps)
        copy_s = time.perf_counter() - t0
        barrier()
        del h_in, d_in, d_o, h_o
    except Exception as e:  # pragma: no cover
        log("copy-ceiling probe failed:", repr(e))

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, e2e_s_u64, e2e_s_u32, copy_s or 0.0, e2e_times.get("packed", 0.0)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_s_u64, e2e_s_u32 = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        copy_s = float(t[4]) or None
        if "packed" in e2e_times:
            e2e_times["packed"] = float(t[5])
    if rank != 0:
        idx.close()
        del d_batches, h_batches, text
        torch.cuda.empty_cache()
        return None

    value = world * args.steps * batch / (total_ms / 1e3)
    e2e_value = world * e2e_steps * batch / e2e_s

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peak()
    line_bytes = int(info.line_bytes)
    t_launch_s = total_ms / args.steps / 1e3
    mean_of = lambda per_batch: float(np.mean([per_batch[i % NB] for i in range(args.steps)]))
    steps_m, lookups_m, checks_m, halves_m, lines_m = (mean_of(x) for x in (steps_per_batch, lookups_per_batch, checks_per_batch,
                                                                            halves_per_batch, lines_per_batch))
    # (1) the contract's model: every executed rank step reads 2 x L lines, a half step 2 lines, every
    #     k-mer table lookup one line, every text verification two (suffix-array entry, text window)
    model_bytes = steps_m * 2 * L * line_bytes + halves_m * 2 * line_bytes + lookups_m * 128 + checks_m * 256
    # (2) what the launch really asked the memory system for: level lines LOADED (sp and ep share one load when they
    #     fall in one line; counted by the instrumented kernel) + the same table / verification lines
    have_lines = lines_m > 0
    executed_bytes = (lines_m * line_bytes + lookups_m * 128 + checks_m * 256) if have_lines else model_bytes
    kernel_name = ("count2_tma_kernel" if os.environ.get("CSFM_PATTERN_STAGING") == "tma" else
                   (("count2q_kernel<false> + count2_kernel<true,false>" if kernels_per_step == 2 else "count2_kernel<true,false>")
                    if int(info.text_check) else "count2_kernel<false,false>")) \
        if int(info.layout) == 2 else ("count3_kernel" if int(info.layout) == 3 else "count_kernel")
    # (3) DRAM bytes of one launch from the committed ncu capture, only if it was taken from THESE kernel sources
    traffic, traffic_meta = committed_traffic(args.workload if not args.large_table else args.workload + "_large_table", kernel_name)
    in_l2 = int(info.blob_bytes) <= L2_RESIDENT_BYTES
    if in_l2:
        l2_peak, l2_src = l2_random_peak(count_working_set(info), line_bytes)  # what the search touches: lines + table
        achieved = executed_bytes / t_launch_s / 1e9
        roofline = {"bound": "l2", "kernel": kernel_name, "achieved": achieved, "peak": l2_peak, "unit": "GB/s",
                    "frac": achieved / l2_peak, "traffic": traffic, "peak_source": l2_src,
                    "count_working_set_bytes": count_working_set(info),
                    "frac_basis": "executed line fetches x line bytes over the measured L2 random-fetch ceiling for a working set of this size",
                    "hbm_stream_peak": peak}
    else:
        basis = "dram traffic (ncu) / duration" if traffic else "executed line fetches x 128 B / duration (no valid ncu capture for these sources)"
        achieved = (traffic if traffic else executed_bytes) / t_launch_s / 1e9
        # the ceiling that binds: dependent random fetches per second out of HBM, for this line size (128-byte lines:
        # profiles/r1_gather_probe.json; 64-byte lines: the 1 GiB row of profiles/r2_l2_sweep_probe.json)
        fetch_ceiling = RANDOM_FETCH_CEILING if line_bytes == 128 else l2_random_peak(1 << 30, line_bytes)[0] * 1e9 / line_bytes
        fetches = (lines_m + lookups_m + 2 * checks_m) if have_lines else executed_bytes / line_bytes
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "frac_basis": basis,
                    "line_fetches_per_s": fetches / t_launch_s,
                    "random_fetch_ceiling_lines_per_s": fetch_ceiling,
                    "frac_of_random_fetch_ceiling": fetches / t_launch_s / fetch_ceiling,
                    "frac_of_random_fetch_ceiling_note": "index fetches the launch executed (level lines + table entries + suffix-array "
                                                         "entries + text windows) per second over the probe's ceiling for this line size"}
        if traffic and line_bytes == 128:
            # the same by DRAM traffic: every 128 bytes moved counted as one line (includes the streamed batch and results)
            roofline["dram_lines_per_s"] = traffic / 128 / t_launch_s
            roofline["frac_of_random_fetch_ceiling_by_dram_traffic"] = traffic / 128 / t_launch_s / fetch_ceiling
    roofline.update({
        "traffic_capture": traffic_meta,
        "model": {"bytes_per_launch": model_bytes, "gbs": model_bytes / t_launch_s / 1e9, "frac_of_hbm_peak": model_bytes / t_launch_s / 1e9 / peak,
                  "note": "SURVEY §8d contract formula: rank steps x 2 x L x line + half steps x 2 x line + lookups x 128 + verifications x 256; "
                          "double-counts the line sp and ep share, so it can exceed what is fetched"},
        "executed": {"bytes_per_launch": executed_bytes, "gbs": executed_bytes / t_launch_s / 1e9,
                     "level_lines_fetched_per_launch": lines_m if have_lines else None},
        "search_steps_per_launch": steps_m, "table_lookups_per_launch": lookups_m, "text_checks_per_launch": checks_m,
        "half_steps_per_launch": halves_m, "half_table": int(info.half_table), "kmer_k": int(info.kmer_k),
        "text_check": int(info.text_check),
        "kernel_ms_mean": float(step_ms.mean()), "kernel_ms_min": float(step_ms.min()), "kernel_ms_max": float(step_ms.max()),
        "kernel_ms_p50": float(np.median(step_ms)),
        "note": "duration per launch = CUDA events on the launching stream around each step (32-byte cursor memset + the count kernel)"})
    # the same launch charged with the REFERENCE algorithm's traffic model: every
    # character of a text-sampled pattern is one executed step of 2 x 8 levels x 64 B there
    ref_steps = float(np.mean([int(d_batches[i % NB][1][-1].item()) for i in range(min(args.steps, NB))]))
    roofline["reference_model"] = {"steps_per_launch": ref_steps, "bytes_per_launch": ref_steps * 1024,
                                   "gbs": ref_steps * 1024 / t_launch_s / 1e9,
                                   "note": "what the reference's step-by-step search over 8 binary levels would move for the same batch "
                                           "(1024 B per step); the engine avoids most of it (16-ary levels, k-mer table, shared lines, "
                                           "text verification), so this is work avoided, not bandwidth achieved"}
    if stepping:
        pb = stepping["search_steps_per_launch"] * 2 * L * line_bytes + stepping["table_lookups_per_launch"] * 128
        pe = stepping["level_lines_fetched_per_launch"] * line_bytes + stepping["table_lookups_per_launch"] * 128
        t2, t2_meta = committed_traffic(args.workload + "_stepping", "count2_kernel<false,false>")
        ts = stepping["ms_per_launch"] / 1e3
        stepping.update({"model_bytes_per_launch": pb, "model_gbs": pb / ts / 1e9, "executed_bytes_per_launch": pe,
                         "executed_gbs": pe / ts / 1e9, "traffic": t2, "traffic_capture": t2_meta,
                         "achieved_gbs": (t2 if t2 else pe) / ts / 1e9, "frac": (t2 if t2 else pe) / ts / 1e9 / peak,
                         "frac_basis": "dram traffic (ncu) / duration" if t2 else "executed line fetches x 128 B / duration",
                         "frac_of_random_fetch_ceiling": (t2 if t2 else pe) / 128 / ts / RANDOM_FETCH_CEILING,
                         "note": "the plain FM-index kernel on the same batches (every character a rank step; selected by asking for "
                                 "the intervals): the robust figure that does not depend on the text being random. An index built with "
                                 "CSFM_BUILD_NO_TEXT_CHECK runs this kernel for every call and carries neither text nor suffix array",
                         "sa_free_index_bytes": int(info.blob_bytes) - 5 * n - 64})
        roofline["stepping_only"] = stepping
    if large:
        if "achieved_gbs" in large:
            t3, t3_meta = committed_traffic(args.workload + "_large_table", "count2_kernel<true,false>")
            tl = large["ms_per_launch"] / 1e3
            pe = large["level_lines_fetched_per_launch"] * line_bytes + large["table_lookups_per_launch"] * 128 + large["text_checks_per_launch"] * 256
            large.update({"executed_bytes_per_launch": pe, "traffic": t3, "traffic_capture": t3_meta,
                          "achieved_gbs": (t3 if t3 else pe) / tl / 1e9, "frac": (t3 if t3 else pe) / tl / 1e9 / peak,
                          "frac_of_random_fetch_ceiling": (t3 if t3 else pe) / 128 / tl / RANDOM_FETCH_CEILING})
        roofline["large_kmer_table"] = large

    # ---- CPU baseline: the unmodified reference on the host cores, bounded sample, checked vs the GPU
    cpu = None
    if world == 1 and not args.no_cpu_baseline and n > (1 << 31):
        cpu = {"value": None, "unit": "queries/s", "kind": "reference", "cores": os.cpu_count(),
               "sample": "not run at this size: the reference's own index (text + BWT + 4n-byte SA + 8 bit planes, 7.2 n bytes) would "
                         "need ~29 GB of host memory injected from the GPU and its rank rescans 0.5 GB per 1-bit: < 0.1 q/s per core "
                         "(extrapolated from c2 / c3, SURVEY §6); parity at this size is tests/test_gpu_fullsize_c5.py"}
    elif world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            if not oracle.ref_available():
                raise RuntimeError("oracle/_ref/libcsref.so missing")
            nthreads = os.cpu_count() or 1
            t0 = time.perf_counter()
            ref = build_reference_index(fm, text, wl)
            t_inject = time.perf_counter() - t0
            hb, ho = h_batches[0]
            data, offs = hb.numpy(), ho.numpy().astype(np.uint64)
            with all_host_cpus():
                out, q, dt = reference_timed_sample(ref, data, offs, args.cpu_budget if full else min(args.cpu_budget, 8.0), nthreads)
            ok = bool((out[:q] == c0[:q].astype(np.uint64)).all())
            cpu = {"value": q / dt, "unit": "queries/s", "cores": nthreads, "kind": "reference",
                   "sample": f"first {q} queries of batch 0 in {dt:.1f} s on {nthreads} std::threads (reference index injected in {t_inject:.0f} s)",
                   "bit_exact_vs_gpu": ok}
            if not ok:
                log("ERROR: GPU counts differ from the reference on the sampled queries")
            del ref
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}

    # independent spot check: a naive scan of the text for a few patterns of batch 0 (no index involved)
    naive_ok = None
    try:
        naive_ok = True
        bytes0, offs0 = d_batches[0]
        for qn in range(4 if n > (1 << 31) else 8):
            pat = bytes0[int(offs0[qn]): int(offs0[qn + 1])]
            m = pat.numel()
            hit = text[: n - m + 1] == pat[0]
            for k in range(1, m):
                hit &= text[k: n - m + 1 + k] == pat[k]
            naive_ok = naive_ok and int(hit.sum()) == int(c0[qn])
            del hit
    except Exception as e:  # pragma: no cover
        naive_ok = f"unavailable: {e!r}"

    e2e = {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "api": "csfm_count_batch_submit_len8/_wait (host pointers, pinned, one length byte per pattern in, u32 counts "
                  "out, 3 steps in flight)",
           "kernels_per_step": 4}
    if "packed" in e2e_times:
        e2e["packed_api"] = {"value": world * e2e_steps * batch / e2e_times["packed"], "ms_per_step": 1e3 * e2e_times["packed"] / e2e_steps,
                             "h2d_bytes_per_step": e2e_bytes["packed"][0], "d2h_bytes_per_step": e2e_bytes["packed"][1],
                             "bits_per_symbol": wire_bits, "kernels_per_step": 5,
                             "api": "csfm_count_batch_submit_packed/_wait (patterns as packed wire codes, one length byte per pattern in, u32 "
                                    "counts out; codes unpacked on the device)",
                             "note": "the copy ceiling below is for the byte form's traffic; this form moves fewer bytes per query"}
    if copy_s:
        ceil_v = world * e2e_steps * batch / copy_s
        e2e["copy_ceiling"] = {"value": ceil_v, "unit": "queries/s", "ms_per_step": 1e3 * copy_s / e2e_steps,
                               "h2d_gbs_per_gpu": h2d * e2e_steps / copy_s / 1e9, "d2h_gbs_per_gpu": d2h * e2e_steps / copy_s / 1e9,
                               "note": "the same H2D + D2H bytes per step from/to pinned memory, three streams per rank like the "
                                       "product's three slots, every rank at once, no kernel: what the host<->device links of this box "
                                       "allow at this N"}
        e2e["frac_of_copy_ceiling"] = e2e_value / ceil_v
    if full:
        e2e.update({
            "u32_api": {"value": world * e2e_steps * batch / e2e_s_u32, "ms_per_step": 1e3 * e2e_s_u32 / e2e_steps,
                        "h2d_bytes_per_step": e2e_bytes["u32"][0], "d2h_bytes_per_step": e2e_bytes["u32"][1],
                        "api": "csfm_count_batch_submit32/_wait (u32 offsets and counts)"},
            "u64_api": {"value": world * e2e_steps * batch / e2e_s_u64, "ms_per_step": 1e3 * e2e_s_u64 / e2e_steps,
                        "h2d_bytes_per_step": h2d64, "d2h_bytes_per_step": d2h64,
                        "api": "csfm_count_batch_submit/_wait (u64 offsets and counts)"},
            "sync_call_ms_per_step": e2e_sync_ms, "sync_call_value": world * batch / (e2e_sync_ms / 1e3)})
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": wl["desc"], "n": n, "levels": L, "line_bytes": int(info.line_bytes), "layout": int(info.layout), "position_samples": int(info.position_samples),
                   "sigma": int(info.sigma), "batch_per_gpu": batch,
                   "distinct_batches": NB, "index_bytes": int(info.blob_bytes), "parallelism": f"dp{world} (index replicated)",
                   "l2_policy": f"inputs larger than L2: {info.blob_bytes / 1e9:.2f} GB index + a different batch every step" if not in_l2
                   else f"index ({info.blob_bytes / 1e6:.0f} MB) is L2-resident at this size; a different batch every step",
                   "index_build_s": build_s, "index_broadcast_ms": bcast_ms, "full_size": n == wl.get("n", 1 << wl["n_log2"]),
                   "build_flags": "CSFM_BUILD_LARGE_TABLE" if args.large_table else "default"},
        "e2e": e2e,
        "gpu_launches": args.steps * kernels_per_step,  # the timed region of `value`: the count kernel(s) of every step
        "gpu_launches_all_timed_legs": launches + e2e_steps * ((1 + 3 + 4) if full else 4),  # + stepping-only, large-table, e2e forms (u64 1, u32 3, len8 4 kernels per step)
        "roofline": roofline,
        "cpu_baseline": cpu,
        "construction": construction,
        "clocks": dict(clocks.summary(), e2e_regions={"sm_mhz": float(np.median(e2e_clock_samples)) if e2e_clock_samples else None,
                                                      "samples": len(e2e_clock_samples), "reasons": sorted(e2e_clock_reasons)}),
        "checks": {"all_counts_ge_1": True, "e2e_equals_device": e2e_equal, "compact_equals_u64_api": compact_equal,
                   "naive_text_scan_equals_counts": naive_ok,
                   "bit_exact_vs_reference_sample": (cpu or {}).get("bit_exact_vs_gpu")},
    }
    idx.close()
    del d_batches, h_batches, text
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------
# configs[0]: the reference's own tools/benchmark.cpp workload (100 KB text, single queries)
# ------------------------------------------------------------------------------------------------
PUBLISHED_C1 = {"random_count_qps": 41530, "frequent_count_qps": 82803, "locate_qps": 8, "build_ms": 4400,
                "source": "reference README.md:135-140 (unspecified Windows x64 host, MSVC Release, 1 thread)"}


def leg_c1(local_rank, cpu_budget):
    """BASELINE.json configs[0]. (a) the reference's tools/benchmark.cpp, compiled UNMODIFIED against the drop-in
    cs::FMIndex (build/ref_callers/bin/benchmark): its own QPS / latency lines, every query one C-ABI call on the
    GPU; (b) the same lists through count_batch / locate_batch (cs_benchmark_batch --batch); (c) the unmodified
    reference on one host core over the same lists. Pattern lists and checksums: tests/golden/c1_workload.npz
    (generated from the compiled reference by tests/golden/make_golden.py)."""
    import re
    import subprocess
    import tempfile
    z = np.load(os.path.join(ROOT, "tests", "golden", "c1_workload.npz"))
    text = z["text"]
    rand = [text[p:p + 5].tobytes() for p in z["rand_pos"]]
    freq10 = [bytes(p)[:l] for p, l in zip(z["freq_patterns"], z["freq_len"])]
    freq = [freq10[i % 10] for i in range(10000)]
    want = {"random": 9907582, "frequent": 16309000, "locate": 163090}
    env = dict(os.environ, CS_DEVICE=str(local_rank))
    out = {"workload": "C1: tools/benchmark.cpp — 100 001-byte text, 10 000 random len-5 + 10 000 frequent count() queries, "
                       "100 locate() queries, one query per call", "published": PUBLISHED_C1}

    def pat_file(d, name, pats):
        path = os.path.join(d, name)
        with open(path, "wb") as f:
            f.write(struct.pack("<I", len(pats)))
            for q in pats:
                f.write(struct.pack("<I", len(q)))
                f.write(q)
        return path

    # (a) the reference's benchmark binary on the drop-in
    exe = os.path.join(ROOT, "build", "ref_callers", "bin", "benchmark")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe], capture_output=True, text=True, errors="replace", timeout=300, env=env)
            txt = r.stdout
            blocks = re.split(r"\n\s*(Random patterns|Frequent patterns|Locate \(locate\)):\n", txt)
            res = {}
            for i in range(1, len(blocks) - 1, 2):
                body = blocks[i + 1]
                g = lambda pat: float(re.search(pat, body).group(1))
                res[blocks[i]] = {"qps": g(r"Throughput:\s+([0-9.]+) QPS"), "p50_us": g(r"Latency p50:\s+([0-9.]+)"),
                                  "p95_us": g(r"Latency p95:\s+([0-9.]+)"), "p99_us": g(r"Latency p99:\s+([0-9.]+)"),
                                  "total_matches": int(g(r"Total matches:\s+([0-9]+)"))}
            bm = re.search(r"Build time:\s+([0-9.]+) ms", txt)
            out["rehosted_benchmark"] = {
                "binary": "build/ref_callers/bin/benchmark (reference tools/benchmark.cpp, unmodified, linked against libcs_b200.so)",
                "random_count": res.get("Random patterns"), "frequent_count": res.get("Frequent patterns"),
                "locate": res.get("Locate (locate)"), "build_ms": float(bm.group(1)) if bm else None,
                "total_matches_equal_reference": bool(res and res["Random patterns"]["total_matches"] == want["random"] and
                                                      res["Frequent patterns"]["total_matches"] == want["frequent"] and
                                                      res["Locate (locate)"]["total_matches"] == want["locate"]),
                "note": "build_ms includes creating the CUDA context (first call of the process)"}
        except Exception as e:  # pragma: no cover
            out["rehosted_benchmark"] = {"unavailable": repr(e)}
    else:
        out["rehosted_benchmark"] = {"unavailable": "build/ref_callers/bin/benchmark not built"}

    # (b) single-query loops and the batch entry points of the drop-in class on the same lists
    tool = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200", "host", "cs_benchmark_batch")
    try:
        with tempfile.TemporaryDirectory() as d:
            tpath = os.path.join(d, "text.bin")
            text.tofile(tpath)
            runs = {}
            for name, cp in (("random", rand), ("frequent", freq)):
                r = subprocess.run([tool, tpath, pat_file(d, name + ".pat", cp), pat_file(d, "loc.pat", freq[:100]), "--batch"],
                                   capture_output=True, text=True, timeout=300, env=env)
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-400:])
                runs[name] = json.loads(r.stdout)
        out["dropin_class"] = {
            "binary": "host/cs_benchmark_batch (host/tools/benchmark_batch.cpp: cs::FMIndex::count / locate per pattern, then count_batch / locate_batch)",
            "build_ms": runs["random"]["build_ms"],
            "random_count_single": runs["random"]["count_single"], "frequent_count_single": runs["frequent"]["count_single"],
            "locate_single": runs["random"]["locate_single"],
            "random_count_batch": runs["random"]["count_batch"], "frequent_count_batch": runs["frequent"]["count_batch"],
            "locate_batch": runs["random"]["locate_batch"],
            "total_matches_equal_reference": bool(runs["random"]["count_single"]["total_matches"] == want["random"] and
                                                  runs["random"]["count_batch"]["total_matches"] == want["random"] and
                                                  runs["frequent"]["count_batch"]["total_matches"] == want["frequent"] and
                                                  runs["random"]["locate_single"]["total_matches"] == want["locate"] and
                                                  runs["random"]["locate_batch"]["occurrences"] == want["locate"])}
        rs = runs["random"]["count_single"]
        out["value"] = rs["qps"]
        out["unit"] = "count queries/s, one query per call (random len-5 list)"
        out["vs_published"] = {"random_count": rs["qps"] / PUBLISHED_C1["random_count_qps"],
                               "frequent_count": runs["frequent"]["count_single"]["qps"] / PUBLISHED_C1["frequent_count_qps"],
                               "locate": runs["random"]["locate_single"]["qps"] / PUBLISHED_C1["locate_qps"]}
    except Exception as e:  # pragma: no cover
        out["dropin_class"] = {"unavailable": repr(e)}
    out["roofline"] = {"bound": "launch latency", "achieved": None, "peak": None, "frac": None, "unit": "us",
                       "note": "a single query is ONE kernel launch (pattern in the kernel parameters, result through mapped pinned "
                               "memory the host spins on) of one warp walking ~6 dependent L2 hits: p50 is launch + PCIe round trip, "
                               "not memory bandwidth; the 0.1 MB index lives in L2. The batch calls move 10 000 queries in one launch"}

    # (c) the unmodified reference on this host, one thread, same lists
    if cpu_budget > 0:
        try:
            import oracle
            ref = oracle.RefIndex(text.tobytes(), stride=32, sa=z["sa"])
            base = {}
            for name, cp in (("random", rand), ("frequent", freq)):
                data, offs = oracle.pack_patterns(cp)
                t0 = time.perf_counter()
                c = ref.count_batch(data, offs, nthreads=1)
                dt = time.perf_counter() - t0
                base[name + "_count_qps"] = len(cp) / dt
                base[name + "_total_matches_ok"] = int(c.sum()) == want[name]
            data, offs = oracle.pack_patterns(freq[:20])
            t0 = time.perf_counter()
            tot, _, _ = ref.locate_batch(data, offs, limit=100000, nthreads=1)
            dt = time.perf_counter() - t0
            base["locate_qps"] = 20 / dt
            base["locate_occ_per_s"] = tot / dt
            out["cpu_baseline"] = {"value": base["random_count_qps"], "unit": "queries/s", "cores": 1, "kind": "reference",
                                   "sample": "full 10 000-query lists for count, 20 of the 100 locate queries, verbatim cs::FMIndex "
                                             "(oracle/_ref) on one host thread like the reference's tool", **base}
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"unavailable: {e!r}"}
    return out


def run_engine(args, rank, world, local_rank):
    """The headline line (the workload named by --workload, default c3 = configs[2]) plus, by default, one leg per
    other BASELINE.json config under `configs`: c1 (rank 0), c2, c4 (= `locate`), c5 — every rank takes part in
    the count / locate legs (weak scaling), rank 0 prints."""
    import torch
    import csfm_b200 as fm
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    affinity = pin_to_gpu_numa_node(local_rank)  # before any pinned allocation: first touch decides the node
    t_start = time.perf_counter()
    line = count_leg(args, args.workload, rank, world, local_rank, args.steps, args.warmup, full=True)
    legs_s = {"main": time.perf_counter() - t_start}
    cpu_ok = not (args.no_cpu_baseline or world > 1)

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            return fn()
        except Exception as e:  # pragma: no cover
            if world > 1:
                raise  # a rank that drops out of the collectives would hang the others
            return {"unavailable": repr(e)}
        finally:
            torch.cuda.empty_cache()
            legs_s[name] = time.perf_counter() - t0

    locate = None
    if not args.no_locate:
        locate = guarded("c4", lambda: measure_locate(fm, dev, rank, world, npat=args.locate_patterns,
                                                     cpu_budget=min(args.cpu_budget, 10.0) if cpu_ok else 0.0))
    configs = {}
    if not args.no_configs and args.workload == "c3" and args.n_log2 is None:
        sub = argparse.Namespace(**vars(args))
        sub.no_cpu_baseline = not cpu_ok
        sub.n_log2 = None
        sub.large_table = False
        c2 = guarded("c2", lambda: count_leg(sub, "c2", rank, world, local_rank, min(args.steps, 100), 5, full=False))
        if rank == 0:
            configs["c1"] = guarded("c1", lambda: leg_c1(local_rank, args.cpu_budget if cpu_ok else 0.0))
        free_gb = torch.cuda.mem_get_info()[0] / 1e9
        if world > 1:
            import torch.distributed as dist
            # every rank must take the same branch below (the leg has collectives): decide on the minimum
            fmin = torch.tensor([free_gb], dtype=torch.float64, device=dev)
            dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
            free_gb = float(fmin[0])
        if free_gb > 150:
            c5_steps = max(3, -(-100 // world))  # the 100 M-pattern sweep of the config, split over the ranks
            c5 = guarded("c5", lambda: count_leg(sub, "c5", rank, world, local_rank, c5_steps, 3, full=False))
        else:
            c5 = {"unavailable": f"needs > 150 GB of free device memory for the 4e9-byte build, {free_gb:.0f} GB free"}
        configs.update({"c2": c2, "c4": "see `locate`", "c5": c5})
    if rank != 0:
        return
    line["config"]["host_affinity"] = affinity
    line["locate"] = locate
    if configs:
        line["configs"] = configs
    line["legs_wall_s"] = legs_s
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n-log2", type=int, default=None, help="override text size (reduced sizes are flagged in config)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--layout", type=int, default=2, choices=[1, 2], help="2 = 16-ary levels / 128-byte lines (default), 1 = binary / 64-byte lines")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--large-table", action="store_true", help="build the MAIN index with CSFM_BUILD_LARGE_TABLE (opt-in mode)")
    ap.add_argument("--no-large-table", action="store_true", help="skip the leg on an index with a 34 GB k-mer table")
    ap.add_argument("--no-locate", action="store_true", help="skip the locate leg (configs[3]: occurrences/s)")
    ap.add_argument("--locate-patterns", type=int, default=1_000_000, help="patterns per GPU in the locate leg (configs[3]: 1 M)")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of the other BASELINE.json configs (c1, c2, c5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_engine(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
