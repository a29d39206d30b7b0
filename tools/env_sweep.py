#!/usr/bin/env python3
"""Runs bench.py once per environment setting and prints one summary line each.

usage: tools/env_sweep.py "CSFM_REFILL_MIN=1" "CSFM_REFILL_MIN=8 CSFM_REFILL_WAIT=4" ... [-- extra bench.py args]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        k = argv.index("--")
        argv, extra = argv[:k], argv[k + 1:]
    for v in argv:
        knob = ""
        env = dict(os.environ)
        for kv in v.split():
            k, _, val = kv.partition("=")
            env[k] = val
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--warmup", "5", "--no-cpu-baseline",
               "--no-locate"] + extra
        p = subprocess.run(cmd, env=env, capture_output=True, text=True)
        try:
            d = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception:
            print(v, "FAILED", p.stderr[-400:])
            continue
        so = d["roofline"].get("stepping_only") or {}
        print(f"[{v}] value={d['value']:.4g} ms={d['ms_per_step']:.4f} e2e={d['e2e']['value']:.4g} "
              f"sync_ms={d['e2e']['sync_call_ms_per_step']:.3f} stepping={so.get('queries_per_s')} checks={d['checks']}", flush=True)


if __name__ == "__main__":
    main()
