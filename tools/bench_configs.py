#!/usr/bin/env python
"""Time the five BASELINE.json configurations (tools/workloads.py): one JSON line per configuration and path.

  python tools/bench_configs.py [--precision bf16|fp32] [--configs 1,2,3,4,5] [--steps 10] [--impl native|eager|trainer]

native   public module API on the native kernels: model -> reference-style loss -> backward -> torch.optim.Adam
eager    the reference's ops in eager PyTorch on the same GPU (the staged reference classes when baseline/_ref exists)
trainer  SirenTrainer: the whole step as one CUDA graph (cfg1..cfg4)
SIREN_PROFILE=1 adds the per-kernel times of the native path.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import workloads  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tasks", type=int, default=8)
    ap.add_argument("--impl", default="native", choices=["native", "eager", "trainer"])
    a = ap.parse_args()
    for c in [int(v) for v in a.configs.split(",")]:
        if a.impl == "trainer":
            if c <= 4:
                print(json.dumps(workloads.run_trainer_config(c, a.precision, steps=a.steps, warmup=a.warmup)), flush=True)
            continue
        print(json.dumps(workloads.run_config(c, a.impl, a.precision, steps=a.steps, warmup=a.warmup, tasks=a.tasks,
                                              want_profile=bool(os.environ.get("SIREN_PROFILE")))), flush=True)
