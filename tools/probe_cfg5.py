#!/usr/bin/env python
"""Developer probe: per-kernel times of cfg5 (8 tasks x 65,536, per-task weights) with the Fourier features
materialised (reference flow) and built inside the kernels (lazy prologue)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads  # noqa: E402

for prec in ("bf16", "fp32"):
    for lazy in (False, True):
        r = workloads.run_config(5, "native", prec, steps=10, warmup=5, want_profile=True, lazy_fourier=lazy)
        print(json.dumps({"precision": prec, "lazy": lazy, "ms": r["ms_per_step"], "kernels": r.get("kernels_us_per_step")}), flush=True)
