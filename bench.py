#!/usr/bin/env python
"""bench.py - headline measurement of the B200 hot path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): Kajita2003 preview control, 4096 random straight/circle footstep walks,
320-tap preview window, 5 ms ticks.  A "step" of the bench is one pass of the hot path over the whole batch
(every preview step of every walk); the metric unit is one preview step = one OneIterationOfPreview call for
both axes (PreviewControl.cpp:324-374).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path through the C ABI
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port) on host cores

With N > 1 (torchrun) every rank runs the same 4096-walk batch shape on its own GPU with its own seed
(instances are independent: weak scaling, no data-path collective); value = all steps of all ranks / max time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Kajita preview steps/s (batched; Herdt2010 QP solves/s reported under 'herdt')"
UNIT = "preview steps/s"
WORKLOAD = "kajita2003_preview_batched_4096_walks_320tap"
FLOP_PER_STEP = 1360.0          # both axes, 2*(2*320 + 40)
FIR_FLOP_PER_STEP = 1280.0      # the 2x320-tap window MACs, the part the FIR kernel executes
BYTES_PER_STEP = 80.0           # streaming minimum: 16 B in, 48 B CoM + 16 B ZMP out


class ClockSampler:
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for (t, line) in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        if not sm:  # region shorter than the sampling period: use the nearest samples
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:] if r[1].split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_preview_rate(offsets, z, seconds=10.0, threads=None):
    """Reference CPU path (oracle port in the reference's deque/AoS layout), one walk per thread."""
    import oracle_lib as ol
    threads = threads or host_cores()
    g = ol.OracleGains(0.005, 1.6, 0.814, 1)
    B = len(offsets) - 1
    # bounded sample: first `nb` walks, grown until the call takes a few seconds
    nb = min(B, max(threads * 2, 32))
    n = int(offsets[-1])
    com = np.zeros((n, 6)); zmp = np.zeros((n, 2))
    best = None
    t_used = 0.0
    while True:
        off = offsets[:nb + 1]
        st = np.zeros((nb, 8))
        t = time.perf_counter()
        _, _, steps = ol.oracle_preview_batch(g, off, z, st, threads=threads, out=(com, zmp))
        dt = time.perf_counter() - t
        t_used += dt
        rate = steps / dt
        best = (rate, steps, nb, dt)
        if dt >= seconds / 3 or nb >= B or t_used > seconds:
            break
        nb = min(B, int(nb * max(2.0, (seconds / 3) / max(dt, 1e-3))))
    return {"value": best[0], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {best[2]} of {B} walks ({best[1]} preview steps, {best[3]:.2f} s), one walk per thread, "
                      "oracle port of PreviewControl::OneIterationOfPreview over std::deque<ZMPPosition>"}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    from jrl_walkgen_b200 import workloads
    offsets, z = workloads.preview_batch(args.walks, seed=0)
    cores = host_cores()
    rates, ms = [], []
    import oracle_lib as ol
    g = ol.OracleGains(0.005, 1.6, 0.814, 1)
    nb = min(args.walks, max(cores * 4, 64))
    n = int(offsets[nb])
    com = np.zeros((n, 6)); zmp = np.zeros((n, 2))
    steps = 0
    for it in range(args.warmup + args.steps):
        st = np.zeros((nb, 8))
        t = time.perf_counter()
        _, _, steps = ol.oracle_preview_batch(g, offsets[:nb + 1], z, st, threads=cores, out=(com, zmp))
        dt = time.perf_counter() - t
        if it >= args.warmup:
            rates.append(steps / dt); ms.append(dt * 1e3)
    value = float(steps * len(ms) / (sum(ms) * 1e-3))
    sample = f"each step = first {nb} of {args.walks} walks ({steps} preview steps), one walk per thread"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "walks": args.walks, "NL": 320, "T": 0.005},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_cuda(args):
    rank, local_rank, world = dist_env()
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import workloads
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = wg.Context(local_rank)
    gains = wg.preview_gains(0.005, 1.6, 0.814, wg.MODE_WITHOUT_INITIALPOS)
    ctx.preview_set_gains(gains)
    offsets, z = workloads.preview_batch(args.walks, seed=1000 * rank)
    B = args.walks
    n = int(offsets[-1])
    plan = ctx.preview_plan(offsets)
    steps_per_pass = int(plan.total_steps)

    # ---- device-resident leg (value) -------------------------------------------------------
    dz = ctx.to_device(z)
    st0 = np.zeros((B, 8))
    ds = ctx.to_device(st0)
    dcom = ctx.alloc(n * 48); dzmp = ctx.alloc(n * 16)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def one_pass():
        ctx._check(ctx.lib.wg_memcpy_h2d(ctx.h, ds.ptr, st0.ctypes.data, st0.nbytes))  # reset 256 KB of states
        plan.run(dz, ds, dcom, dzmp, True, mem=wg.WG_MEM_DEVICE)

    for _ in range(max(args.warmup, 3)):
        one_pass()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ctx.reset_launches()
    ctx.prof_begin(4 * args.steps + 8)
    barrier()
    t0 = time.time()
    ctx.timer_start()
    for _ in range(args.steps):
        one_pass()
    ms_total = ctx.timer_stop_ms()
    t1 = time.time()
    barrier()
    launches = ctx.launches
    prof = ctx.prof_end()

    # ---- end-to-end leg: host buffers through the C ABI (H2D + kernels + D2H inside) ---------
    zp = ctx.pinned(z.shape); zp[:] = z
    stp = ctx.pinned((B, 8))
    comp = ctx.pinned((n, 6)); zmpp = ctx.pinned((n, 2))
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        stp[:] = 0.0
        plan.run(zp, stp, comp, zmpp, True, mem=wg.WG_MEM_HOST)
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        stp[:] = 0.0
        plan.run(zp, stp, comp, zmpp, True, mem=wg.WG_MEM_HOST)
    ctx.sync()
    e2e_s = time.perf_counter() - te
    clocks = sampler.stop(t0, t1)
    h2d = z.nbytes + stp.nbytes
    d2h = comp.nbytes + zmpp.nbytes + stp.nbytes

    # ---- max over ranks ----------------------------------------------------------------------
    total_steps_all = steps_per_pass
    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        s = torch.tensor([steps_per_pass], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        total_steps_all = float(s[0])
    ms_per_step = ms_total / args.steps
    value = total_steps_all / (ms_per_step * 1e-3)
    e2e_value = total_steps_all * e2e_steps / e2e_s

    if rank == 0:
        # roofline of the dominant kernel (largest share of the timed region)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = ctx.fp64_peak_tflops()
        names = {6: "preview_fused_kernel", 2: "herdt_qp_kernel", 3: "herdt_mpc_kernel", 4: "pldp_kernel"}
        kern = {names.get(k, str(k)): {"launches": v[0], "avg_ms": v[1] / v[0]} for k, v in prof.items()}
        dom = max(kern.items(), key=lambda kv: kv[1]["avg_ms"] * kv[1]["launches"])
        dom_name, dom_ms = dom[0], dom[1]["avg_ms"]
        if dom_name == "preview_recur_kernel":
            # 4-state recursion: streams fir(16 B) + zmpref(16 B) in, CoM(48 B) + ZMP(16 B) out per step
            ach = 96.0 * steps_per_pass / (dom_ms * 1e-3) / 1e9
            roof = {"kernel": dom_name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src,
                    "algorithmic_bytes_per_step": 96.0}
        else:
            fl = (FIR_FLOP_PER_STEP if dom_name == "preview_fir_kernel" else FLOP_PER_STEP)
            ach = fl * steps_per_pass / (dom_ms * 1e-3) / 1e12
            roof = {"kernel": dom_name, "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak, "traffic": None,
                    "peak_source": "measured live: register-resident DFMA chains on all SMs (wg_measure_fp64_peak); "
                                   "MEASURED_PEAKS.json has no FP64 entry",
                    "algorithmic_flop_per_step": fl}
        roof["hbm_frac_streaming_minimum"] = BYTES_PER_STEP * steps_per_pass / (ms_per_step * 1e-3) / 1e9 / hbm_peak
        cpu = None
        if world == 1 or True:
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
            cpu = cpu_preview_rate(offsets, z, seconds=args.cpu_seconds)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "walks_per_gpu": B, "NL": 320, "T": 0.005,
                           "preview_steps_per_pass_per_gpu": steps_per_pass,
                           "l2": "inputs+outputs per pass (%.2f GB) exceed the 126 MB L2" % ((n * 80) / 1e9)},
                "roofline": roof, "kernels": kern, "fp64_peak_tflops_measured": fp64_peak,
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "passes": e2e_steps, "api": "wg_preview_run_batch(WG_MEM_HOST), pinned host buffers"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    plan.destroy()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--walks", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
