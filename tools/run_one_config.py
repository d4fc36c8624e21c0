#!/usr/bin/env python
"""Developer aid: ONE configuration for a few steps (for `ncu -k regex:... python tools/run_one_config.py 3 bf16`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
r = workloads.run_config(cfg, "native", prec, steps=1, warmup=1)
print(cfg, prec, round(r["ms_per_step"], 3), "ms")
