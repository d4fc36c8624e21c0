#!/usr/bin/env python
"""bench.py -- SIREN train-step throughput (coords/sec, fwd + bwd + Adam) on N B200s.

Workload (BASELINE.json configs[1], "cfg2"): SIREN 3x256, in=2, out=1, full-batch 512x512 =
262,144 coordinates per step, synthetic image, MSE (image_mse high_freq=False), Adam.
A "step" = one pass of the hot path over that batch: forward, loss gradient, backward (dgrad +
wgrad), gradient all-reduce (N > 1) and the fused Adam update, replayed as one CUDA graph.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--scaling weak|strong]
  python bench.py --impl reference ...    # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  Multi-GPU: launched by torch.distributed.run, one rank per GPU.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIDE = 512
N_COORDS = SIDE * SIDE
D_IN, D_OUT, HIDDEN, N_HIDDEN = 2, 1, 256, 3
U = 2 * (D_IN * HIDDEN + N_HIDDEN * HIDDEN * HIDDEN + HIDDEN * D_OUT)      # forward FLOP / coordinate
FLOP_PER_COORD = 3 * U                                                      # fwd + dgrad + wgrad
HIDDEN_LAYER_FLOP = 2 * HIDDEN * HIDDEN                                     # one 256x256 layer, per coordinate


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (NVML, ~2 ms per sample; falls back to
    polling nvidia-smi when the NVML binding is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-style remapping when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                parts = [v.strip() for v in vis.split(",")]
                if index < len(parts) and parts[index].isdigit():
                    phys = int(parts[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
        self.samples.append((sm, mx, flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        parts = [p.strip() for p in out.stdout.strip().split(",")]
        if len(parts) >= 6 and parts[0].isdigit() and parts[1].isdigit():
            self.samples.append((int(parts[0]), int(parts[1]), [p.lower().startswith("active") for p in parts[2:6]]))

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=10)
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if s[2][i]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import siren_ref_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # bounded sample: one probe step at full size; if K + W full steps would not end within a few
    # minutes, every step processes a proportionally smaller slice of the 262144 coordinates
    probe = siren_ref_port.time_steps(N_COORDS, steps=1, warmup=0, threads=threads)
    budget_s = 150.0
    total = (args.steps + max(args.warmup, 1)) * probe
    sample_n = N_COORDS
    if total > budget_s:
        sample_n = max(8192, int(N_COORDS * budget_s / total) // 1024 * 1024)
    sec = siren_ref_port.time_steps(sample_n, steps=args.steps, warmup=max(args.warmup, 1), threads=threads)
    value = sample_n / sec
    sample_txt = ("full 262144-coord step" if sample_n == N_COORDS else
                  "%d-coord slice of the 262144-coord step (bounded run time)" % sample_n)
    line = {
        "impl": "reference", "metric": "siren_train_step_coords_per_sec", "value": value, "unit": "coords/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg2: SIREN 3x256 image fit, 512x512 = 262144 coords/step, MSE + Adam",
                   "coords_per_step": N_COORDS, "coords_per_timed_step": sample_n},
        "cpu_baseline": {"value": value, "unit": "coords/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%s x %d (torch CPU ops restating modules.py:25-26,38 + autograd + Adam)"
                                   % (sample_txt, args.steps)},
        "e2e": {"value": value, "unit": "coords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 5 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 200 if args.steps is None else args.steps
    args.warmup = 10 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from siren_mri_b200 import _lib, modules
    from siren_mri_b200.trainer import SirenTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.siren_b200_device_ok(), "siren_b200_device_ok")

    # per-GPU shard of the coordinate batch
    if args.scaling == "weak":
        n_local, n_global = N_COORDS, N_COORDS * world
    else:
        n_global = N_COORDS
        n_local = (N_COORDS + world - 1) // world

    torch.manual_seed(0)
    model = modules.SingleBVPNet(in_features=D_IN, out_features=D_OUT, hidden_features=HIDDEN,
                                 num_hidden_layers=N_HIDDEN, precision=args.precision).to(dev)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    trainer = SirenTrainer(model, n_local, lr=1e-4, loss_weight=1.0 / 16384.0, precision=args.precision,
                           use_graph=not args.no_graph)

    # synthetic data of the config's shape: a 512x512 grid in [-1,1]^2 (per rank: its shard of a
    # 512 x (512*world) strip under weak scaling) and a smooth synthetic image in [-1,1]
    # (the image is ONE function of the coordinates, the same on every rank, so the ranks fit one consistent scene)
    g = torch.Generator().manual_seed(1234)
    lin = torch.linspace(-1, 1, SIDE)
    lin_x = lin
    if args.scaling == "weak" and world > 1:      # this rank's 512 columns of the strip, the strip scaled to [-1, 1]
        lin_x = torch.linspace(-1.0 + 2.0 * rank / world, -1.0 + 2.0 * (rank + 1) / world, SIDE + 1)[:-1]
    grid = torch.stack(torch.meshgrid(lin, lin_x, indexing="ij"), dim=-1).reshape(1, -1, 2)
    if n_local != N_COORDS:
        from siren_mri_b200.parallel import shard_bounds
        b, e = shard_bounds(N_COORDS, rank, world)
        grid = grid[:, b:e]
        if grid.shape[1] < n_local:
            grid = torch.cat([grid, grid[:, :n_local - grid.shape[1]]], dim=1)
    img = torch.zeros(1, grid.shape[1], 1)
    for _ in range(8):
        f = torch.randn(2, generator=g) * 6.0
        ph = torch.rand(1, generator=g) * 6.28
        img += torch.sin(grid @ f.view(2, 1) + ph)
    img = img / 8.0           # |sum of 8 sines| <= 8: in [-1, 1] on every rank with one scale
    coords_host = grid.contiguous().pin_memory()
    gt_host = img.contiguous().pin_memory()
    trainer.coords.copy_(coords_host)
    trainer.gt.copy_(gt_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def log(msg):
        if rank == 0:
            print("[bench] %s" % msg, file=sys.stderr, flush=True)

    # ---------------- device-resident timing: W warm-up, K timed steps ----------------
    log("warm-up (world=%d, n_local=%d)" % (world, n_local))
    for _ in range(args.warmup):
        trainer.step()
    barrier()
    log("timed region")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        trainer.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_global / (ms_step * 1e-3)
    loss_after = float(trainer.loss.item())

    # ---------------- end-to-end through the public API with host buffers ----------------
    log("end-to-end")
    # every step uploads its batch (pinned host memory -> device) and its loss is read back on the host, all inside
    # the timed region; the public call is the pipelined one: the loss of step k is read after step k+1 has been
    # submitted, so the upload of k+1 runs under the kernels of k (SirenTrainer.submit_from_host)
    e2e_steps = max(5, min(args.steps, 100))
    for _ in range(8):
        trainer.submit_from_host(coords_host, gt_host).result()
    barrier()
    ev0.record()
    prev, e2e_loss = None, None
    for _ in range(e2e_steps):
        h = trainer.submit_from_host(coords_host, gt_host)
        if prev is not None:
            e2e_loss = prev.result()
        prev = h
    e2e_loss = prev.result()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e = {"value": n_global / (e2e_ms * 1e-3), "unit": "coords/s",
           "h2d_bytes_per_step": int(coords_host.numel() * 4 + gt_host.numel() * 4) * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms,
           "api": "SirenTrainer.submit_from_host (loss read one step late)", "last_loss": e2e_loss}

    # ---------------- per-kernel timing (CUDA events around each launch, ungraphed) ----------------
    # every rank runs the same steps (they contain the gradient all-reduce); rank 0 records events
    roofline, kernel_table = None, None
    launches_per_step = trainer.kernels_per_step          # replaced below by the count the library itself reports
    log("per-kernel timing")
    saved_graph, saved_flag = trainer.graph, trainer.use_graph
    trainer.use_graph = False
    prof_steps = 10
    trainer.step()
    barrier()
    if rank == 0:
        lib.siren_b200_profile_begin()
    for _ in range(prof_steps):
        trainer.step()
    barrier()
    trainer.use_graph, trainer.graph = saved_flag, saved_graph
    if rank == 0:
        pk = peaks()
        buf = ctypes.create_string_buffer(1 << 16)
        lib.siren_b200_profile_end(buf, len(buf))
        kernel_table = {}
        for ln in buf.value.decode().strip().splitlines():
            name, cnt, ms = ln.split()
            kernel_table[name] = {"launches": int(cnt), "avg_us": 1e3 * float(ms) / int(cnt),
                                  "us_per_step": 1e3 * float(ms) / prof_steps}
        # one entry per launch site of the library; the "adam" and "clip_grad" sites launch two kernels each
        launches_per_step = (sum(v["launches"] for v in kernel_table.values()) +
                             kernel_table.get("adam", {}).get("launches", 0) +
                             kernel_table.get("clip_grad", {}).get("launches", 0)) // prof_steps
        # The tensor-core kernels of this per-layer design are HBM-bound (87 FLOP/B against a machine
        # balance of 253 FLOP/B): the roofline of the dominant kernel is reported
        # against the measured copy bandwidth, with its tensor-pipe numbers next to it.
        #   algorithmic bytes per coordinate and launch (bf16 planes of 256 features = 512 B):
        #   hidden_fwd  read h (512) + write h', c' (1024); hidden_dgrad read zbar, c (1024) + write zbar' (512)
        #   wgrad       read zbar_l, h_{l-1} (1024) per hidden layer; fp32-parity mode doubles every plane
        #   mlp_fused_fwd (bf16 mode): the whole forward in one launch -- coordinates in (4 d), ONE fp16 phase plane
        #               per sine layer out ((N_HIDDEN + 1) x 512), y out (4 o); activations never travel as operands
        #   mlp_fused_bwd: the dgrad chain from the loss gradient down -- gy and coordinates in, the phase plane of
        #               every sine layer in ((N_HIDDEN + 1) x 512), the adjoints of sine layers N_HIDDEN .. 1 out
        #               (N_HIDDEN x 512; they are the weight-gradient kernel's operands)
        pf = 1 if args.precision == "bf16" else 2
        abytes = {"hidden_fwd": 1536 * pf * n_local, "hidden_dgrad": 1536 * pf * n_local,
                  "wgrad": (1024 * pf * N_HIDDEN - (512 if args.precision == "bf16" else 0)) * n_local,
                  "mlp_fused_fwd": (N_HIDDEN * 512 + 4 * D_IN + 4 * D_OUT) * n_local,
                  "mlp_fused_bwd": (2 * N_HIDDEN * 512 + 4 * D_IN + 4 * D_OUT) * n_local}
        flops = {"hidden_fwd": HIDDEN_LAYER_FLOP * n_local, "hidden_dgrad": HIDDEN_LAYER_FLOP * n_local,
                 "wgrad": HIDDEN_LAYER_FLOP * n_local * N_HIDDEN,
                 "mlp_fused_fwd": HIDDEN_LAYER_FLOP * n_local * N_HIDDEN,
                 "mlp_fused_bwd": HIDDEN_LAYER_FLOP * n_local * N_HIDDEN}
        tc = {k: v for k, v in kernel_table.items() if k in flops}
        top = max(tc, key=lambda k: tc[k]["us_per_step"])
        sec = tc[top]["avg_us"] * 1e-6
        achieved = abytes[top] / sec / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        # (profiles/r01_ncu_full_summary.txt, bf16 mode, 262144 coords); None when not captured
        ncu_traffic = {"mlp_fused_fwd": 497.1e6, "mlp_fused_bwd": 954.9e6, "wgrad": 810.6e6,
                       "hidden_fwd": 345.4e6, "hidden_dgrad": 368.2e6}       # last two: SIREN_FUSED_*=0 path
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / pk["hbm_gbs"],
                    "traffic": ncu_traffic.get(top) if (args.precision == "bf16" and n_local == N_COORDS) else None,
                    "algorithmic_bytes_per_launch": abytes[top],
                    "peak_source": pk["source"] + " hbm_gbs (copy)", "avg_us": tc[top]["avg_us"],
                    "kernel_share_of_step": tc[top]["us_per_step"] / sum(v["us_per_step"] for v in kernel_table.values()),
                    "tensor": {"achieved_tflops": flops[top] / sec / 1e12, "peak_tflops": pk["bf16_tflops"],
                               "frac": flops[top] / sec / 1e12 / pk["bf16_tflops"]},
                    "step_frac_of_peak": FLOP_PER_COORD * value / world / 1e12 / pk["bf16_tflops"],
                    "step_frac_of_sustained_peak": (FLOP_PER_COORD * value / world / 1e12 /
                                                    pk["bf16_tflops_sustained"]) if pk["bf16_tflops_sustained"] else None}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import siren_ref_port
        threads = os.cpu_count() or 1
        sec = siren_ref_port.time_steps(N_COORDS, steps=2, warmup=1, threads=threads)
        cpu_baseline = {"value": N_COORDS / sec, "unit": "coords/s", "cores": threads, "kind": "port",
                        "sample": "2 timed full 262144-coord steps after 1 warm-up (oracle/siren_ref_port.py)",
                        "ms_per_step": sec * 1e3}

    if rank == 0:
        line = {
            "metric": "siren_train_step_coords_per_sec", "value": value, "unit": "coords/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3",
            "data": "synthetic",
            "config": {"workload": "cfg2: SIREN 3x256 image fit, 512x512 = 262144 coords/step%s, MSE + Adam"
                                   % (" per GPU" if args.scaling == "weak" and world > 1 else ""),
                       "coords_per_step_global": n_global, "coords_per_gpu": n_local,
                       "precision_mode": args.precision, "parallelism": "coords-dp%d" % world,
                       "l2": "per-step working set (~1.6 GB of activation/stash planes) exceeds the 126 MB L2",
                       "cuda_graph": not args.no_graph, "flop_per_coord": FLOP_PER_COORD},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernel_table,
            "loss_after": loss_after,
        }
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        # leave without tearing the communicator down: destroy_process_group() has been seen to hang
        # when collectives were captured into CUDA graphs; the processes are done anyway
        barrier()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
