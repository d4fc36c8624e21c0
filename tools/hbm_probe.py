#!/usr/bin/env python
"""HBM ceilings for the store-dominated kernels: device-to-device copy (read + write) next to a pure fill
(write only) of the same size.  The fused forward writes 8 planes and reads almost nothing, so the fill
number is the roof it actually sits under."""
import json

import torch


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


if __name__ == "__main__":
    n = 1 << 30                                    # 1 GiB per buffer, far beyond the 126 MB L2
    a = torch.empty(n, dtype=torch.uint8, device="cuda")
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    t_copy = timed(lambda: b.copy_(a))
    t_fill = timed(lambda: a.zero_())
    t_read = timed(lambda: a.view(torch.int32).sum())
    print(json.dumps({"bytes": n, "copy_GBps_rw": 2 * n / t_copy / 1e9, "fill_GBps_w": n / t_fill / 1e9,
                      "reduce_GBps_r": n / t_read / 1e9}))
