#!/usr/bin/env python
"""Developer probe: per-kernel times of cfg3 / cfg4 (public path, bf16 and fp32-parity)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads  # noqa: E402

for cfg in (3, 4):
    for prec in ("bf16", "fp32"):
        r = workloads.run_config(cfg, "native", prec, steps=5, warmup=3, want_profile=True)
        print(json.dumps({"cfg": cfg, "precision": prec, "ms": round(r["ms_per_step"], 3), "kernels": r.get("kernels_us_per_step")}), flush=True)
