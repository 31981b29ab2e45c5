#!/usr/bin/env python
"""Benchmark of the SMC hot path: pfNonlinBS bootstrap particle filter, 2^24 particles x 100 steps,
systematic resampling every step (BASELINE.json metric: particle-steps/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

One bench "step" = one complete filter (Initialise + 99 Iterate + the moment passes) over one synthetic
observation series.  Prints ONE JSON line on rank 0.  For N > 1 launch with torch.distributed.run: the ranks hold the
islands (2^24 particles each) of ONE population of N x 2^24 particles with GLOBAL systematic resampling; the per-step
exchange (three tiny all-gathers + remote parent reads) is done by the kernels over NVLink peer memory (weak scaling).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_N = 24
T_STEPS = 100
ALG_BYTES_PER_PSTEP = 56.0      # B_alg(d) = 16 d + 40, d = 1
ALG_BYTES_MOVE = 28.0           # K1 gather + propagate + weight: 16 d + 12
ALG_BYTES_SCAN = 28.0           # K2 + K3 normalise/scan + ancestor search: 16 + 12
SYSTEMATIC = 3
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback


def synthetic_data(T, seed=20251019):
    """Restated R/simNonlin.R:1-14 (defaults var_init=10, var_evol=10, var_obs=1, cosSeqOffset=-1)."""
    rng = np.random.default_rng(seed)
    zs, zo = rng.standard_normal(T), rng.standard_normal(T)
    x = np.empty(T)
    x[0] = zs[0] * np.sqrt(10.0)
    for i in range(2, T + 1):
        p = x[i - 2]
        x[i - 1] = 0.5 * p + 25.0 * p / (1.0 + p * p) + 8.0 * np.cos(1.2 * (i - 1.0)) + zs[i - 1] * np.sqrt(10.0)
    return x * x / 20.0 + zo


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [c.strip() for c in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


# ---- CPU baseline: the reference's own smc::sampler compiled unmodified (oracle/_ref), else the C port ------
def cpu_lib():
    ref = os.path.join(ROOT, "oracle", "_ref", "libsmcref.so")
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if os.path.exists(ref):
        L = ctypes.CDLL(ref)
        L.smcref_time_pfnonlinbs.restype = ctypes.c_double
        return "reference", L
    if not os.path.exists(port):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), port], check=True, capture_output=True)
    L = ctypes.CDLL(port)
    L.smco_time_pfnonlinbs.restype = ctypes.c_double
    return "port", L


def cpu_time_filter(kind, L, y, n, t):
    yy = np.ascontiguousarray(y[:t], dtype=np.float64)
    yp = yy.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if kind == "reference":
        return L.smcref_time_pfnonlinbs(yp, ctypes.c_long(t), ctypes.c_long(n), SYSTEMATIC, ctypes.c_double(1.01 * n))
    return L.smco_time_pfnonlinbs(yp, ctypes.c_long(t), ctypes.c_long(n), SYSTEMATIC, ctypes.c_double(1.01 * n),
                                  ctypes.c_ulonglong(1))


def host_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"host_cpus": os.cpu_count(), "host_cpu_model": model}


def cpu_baseline(y, log2n=LOG2_N, t=2):
    kind, L = cpu_lib()
    n = 1 << log2n
    sec = cpu_time_filter(kind, L, y, n, t)
    return dict({"value": n * t / sec, "unit": "particle-steps/s", "cores": 1, "kind": kind,
            "sample": f"pfNonlinBS 2^{log2n} particles x {t} steps, systematic resampling every step, 1 thread "
                      f"({'reference smc::sampler compiled from its own sources, shim RNG' if kind == 'reference' else 'C port of the reference path'}), {sec:.1f} s"},
                **host_info())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    y = synthetic_data(T_STEPS)
    kind, L = cpu_lib()
    total = max(1, args.steps + args.warmup)
    t_s, log2n = 2, LOG2_N
    while log2n > 12 and total * (1 << log2n) * t_s / 2.0e6 > 170.0:
        log2n -= 1
    n = 1 << log2n
    for _ in range(args.warmup):
        cpu_time_filter(kind, L, y, n, t_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_time_filter(kind, L, y, n, t_s)
    sec = time.perf_counter() - t0
    value = args.steps * n * t_s / sec
    sample = (f"each step = reference pfNonlinBS filter on 2^{log2n} particles x {t_s} time steps (bounded sample of the "
              f"2^{LOG2_N} x {T_STEPS} workload), systematic, 1 thread")
    line = {"impl": "reference", "metric": "particle-steps/sec (pfNonlinBS, 2^24 particles)", "value": value,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (restated simNonlin, fixed seed)",
            "config": dict(workload_config(args.gpus), reference_sample={"particles": n, "time_steps": t_s,
                           "note": "the CPU path is linear in particles x steps; the rate is per particle-step"}),
            "cpu_baseline": dict({"value": value, "unit": "particle-steps/s", "cores": 1, "kind": kind, "sample": sample}, **host_info()),
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(gpus):
    return {"workload": f"pfNonlinBS bootstrap particle filter, 2^{LOG2_N} particles x {T_STEPS} steps per GPU, SYSTEMATIC resampling every "
                        f"step (threshold 1.01 N), HistoryType NONE, outputs mean/sd/logNC/ESS per step",
            "particles_per_gpu": 1 << LOG2_N, "time_steps": T_STEPS, "resampling": "systematic",
            "parallelism": "1 GPU" if gpus == 1 else f"island partition: one population of {gpus} x 2^{LOG2_N} particles, global systematic "
                           f"resampling, per-step all-gathers and remote parent reads inside the kernels over NVLink peer memory (CUDA IPC)",
            "l2": "working set per step (>= 400 MB of state, weights and ancestor lists) exceeds the 126 MB L2; no explicit flush"}


def verify_island(smc, dist, torch, rank, world, local, stream, log2n_total=20, T=10):
    """Before anything is timed: a short island run (2^20 particles in total x 10 steps) must reproduce the single-GPU run of
    the same population -- traces to 1e-9 relative and the last step's ancestor indices bit for bit."""
    from rcppsmc_b200 import island
    n_glob = 1 << log2n_total
    n_loc = n_glob // world
    y = synthetic_data(T, seed=7)
    f = smc.PfNonlinBS(n_loc, T, seed=4242, stream=stream.cuda_stream, island=(rank, world))
    island.connect(f, device=torch.device("cuda", local))
    f.set_data(y)
    f.run(SYSTEMATIC)
    f.sync()
    out = f.read()
    anc = torch.from_numpy(f.read_ancestors().astype(np.int64)).cuda()
    parts = [torch.empty_like(anc) for _ in range(world)]
    dist.all_gather(parts, anc)
    res = torch.zeros(6, dtype=torch.float64, device="cuda")
    if rank == 0:
        g = smc.PfNonlinBS(n_glob, T, seed=4242, stream=stream.cuda_stream)
        g.set_data(y)
        g.run(SYSTEMATIC)
        g.sync()
        ref = g.read()
        ref_anc = g.read_ancestors().astype(np.int64)
        g.close()
        err = [float(np.max(np.abs(out[k] - ref[k]) / np.maximum(1e-12, np.abs(ref[k])))) for k in ("logNC", "ESS", "mean", "sd")]
        anc_ok = bool(np.array_equal(torch.cat(parts).cpu().numpy(), ref_anc))
        ok = anc_ok and max(err) <= 1e-9 and bool(np.array_equal(out["resampled"], ref["resampled"]))
        res = torch.tensor([1.0 if ok else 0.0, 1.0 if anc_ok else 0.0] + err, dtype=torch.float64, device="cuda")
    dist.broadcast(res, 0)
    f.close()
    r = res.cpu().numpy()
    return {"ok": bool(r[0] == 1.0), "ancestors_bit_exact": bool(r[1] == 1.0), "particles_total": n_glob, "time_steps": T,
            "max_rel_err": {"logNC": float(r[2]), "ESS": float(r[3]), "mean": float(r[4]), "sd": float(r[5])}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (debugging only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import rcppsmc_b200 as smc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rcppsmc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def trace(msg):
        if os.environ.get("SMCB_BENCH_TRACE"):
            sys.stderr.write(f"[rank {rank}] {msg}\n"); sys.stderr.flush()

    n = 1 << args.log2n
    y = synthetic_data(T_STEPS)
    stream = torch.cuda.current_stream()
    if world == 1:
        f = smc.PfNonlinBS(n, T_STEPS, seed=1000, stream=stream.cuda_stream)
    else:
        from rcppsmc_b200 import island
        f = smc.PfNonlinBS(n, T_STEPS, seed=1000, stream=stream.cuda_stream, island=(rank, world))
        island.connect(f, device=torch.device("cuda", local))   # 64-byte IPC handles over the NCCL group: plumbing only
    trace("handles connected")
    island_check = None
    if world > 1:
        island_check = verify_island(smc, dist, torch, rank, world, local, stream)
        trace(f"island check {island_check}")
        if not island_check["ok"]:
            if rank == 0:
                print(json.dumps({"error": "island run does not reproduce the single-GPU run", "check": island_check}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            return 3
    f.set_data(y)
    f.set_profile(False)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): data and tables already in HBM
    for _ in range(args.warmup):
        f.run(SYSTEMATIC)
    trace("warmup enqueued")
    barrier()
    trace("warmup done")
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    move_ms = scan_ms = 0.0
    move_n = scan_n = 0
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        f.run(SYSTEMATIC)
    e1.record(stream)
    barrier()
    trace("timed region done")
    ms = e0.elapsed_time(e1)
    launches = f.launches() * args.steps
    out_dev = f.read()
    if world == 1:   # kernel split: one extra, untimed filter with an event after every launch
        f.set_profile(True)
        f.run(SYSTEMATIC)
        f.sync()
        move_ms, move_n, scan_ms, scan_n = f.kernel_ms()

    # ---- end to end through the public call: host buffers in, host buffers out, every step
    f.set_profile(False)
    yh = np.ascontiguousarray(y)
    if world == 1:
        smc.pfNonlinBS(yh, n, resample_mode=SYSTEMATIC, seed=7, full=True)  # builds the cached engine of the one-shot call
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        if world == 1:
            out_e2e = smc.pfNonlinBS(yh, n, resample_mode=SYSTEMATIC, seed=2000 + k, full=True)
        else:   # island handle: observations host -> device, filter, traces device -> host
            f.set_data(yh)
            f.run(SYSTEMATIC)
            out_e2e = f.read()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    trace("e2e done")
    clk = clocks.stop()

    t_ms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t_ms[0]), float(t_ms[1])
    total_psteps = float(world) * args.steps * n * T_STEPS
    value = total_psteps / (ms_max * 1e-3)
    e2e_value = total_psteps / (e2e_ms_max * 1e-3)

    if rank == 0:
        peak, peak_src = hbm_peak()
        if world == 1:
            dominant = "k_step (gather + propagate + weight)" if move_ms >= scan_ms else "resampling passes (k_rs_mass + k_rs_count + k_rs_map)"
            per_launch_ms = (move_ms / max(1, move_n)) if move_ms >= scan_ms else (scan_ms / max(1, scan_n))
            alg_bytes = (ALG_BYTES_MOVE if move_ms >= scan_ms else ALG_BYTES_SCAN) * n
            tkey = "k_step" if move_ms >= scan_ms else "k_rs"
        else:   # per-kernel events are a single-GPU diagnostic: report the whole step per GPU instead
            dominant = "whole step per GPU (k_step + k_rs_mass + k_rs_count + k_rs_map)"
            per_launch_ms = ms_max / args.steps / T_STEPS
            alg_bytes = ALG_BYTES_PER_PSTEP * n
            tkey = None
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic = profiled_traffic()
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (traffic or {}).get(tkey) if tkey else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": per_launch_ms,
                    "step_share": {"k_step_ms": move_ms, "resampling_passes_ms": scan_ms, "filter_ms": ms_max / args.steps,
                                   "note": "per-launch CUDA events of one extra, untimed filter"},
                    "whole_path": {"achieved": value / world * ALG_BYTES_PER_PSTEP / 1e9, "frac": value / world * ALG_BYTES_PER_PSTEP / 1e9 / peak,
                                   "bytes_per_particle_step": ALG_BYTES_PER_PSTEP}}
        cpu = None if args.no_cpu else cpu_baseline(y)
        line = {"metric": "particle-steps/sec (pfNonlinBS, 2^24 particles)", "value": value, "unit": "particle-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (restated simNonlin, fixed seed); particles drawn on device by Philox4x32-10",
                "config": workload_config(world), "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(2 * 8 * T_STEPS),
                        "d2h_bytes_per_step": int((4 * 8 + 4) * T_STEPS),
                        "note": ("smcb200_pfNonlinBS one-shot C-ABI call with host buffers (observations in, mean/sd/logNC/ESS/resampled out)" if world == 1 else
                                 "island handle, per step: smcb200_pfNonlinBS_set_data (observations host -> device), _run, _read (traces device -> host)")},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "check": {"logNC_final": float(out_dev["logNC"][-1]), "mean_final": float(out_dev["mean"][-1]),
                          "e2e_logNC_final": float(out_e2e["logNC"][-1]), "island_vs_single_gpu": island_check}}
        print(json.dumps(line), flush=True)
    f.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
