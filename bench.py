#!/usr/bin/env python
"""Benchmark of the Morse-wavelet CWT hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo, N ranks via torchrun
    python bench.py --impl reference --steps K --warmup W    # CPU oracle port on the host cores

A step is one pass of the transform over one batch of synthetic chirp + pink-noise
channels.  Workload at N = 1 is config 2 of BASELINE.json (64 channels x 1.25 kHz x
30 min, 96 scales, fp32 amplitude); at N > 1 the channels are sharded over the ranks with
the same 64 channels per GPU (weak scaling, no data-path collective).

One JSON line is printed by rank 0 (see the driver contract in the task statement).
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

METRIC = "cwt_output_coeffs_per_sec"
UNIT = "coeff/s"

WORKLOADS = {
    # name: fs, samples, channels per GPU, freq_limits, voices/octave, output
    "cfg2": dict(fs=1250.0, n=2250000, channels=64, freq_limits=[0.40, 300.0], vpo=10, output="amplitude",
                 desc="64ch x 1.25kHz x 30min, 96 scales, fp32 amplitude"),
    "cfg1": dict(fs=1000.0, n=60000, channels=1, freq_limits=None, vpo=10, output="amplitude",
                 desc="1ch x 1kHz x 60s, 84 scales, fp32 amplitude"),
    # config 3 of BASELINE.json is 256 channels over 8 GPUs = 32 channels per GPU; its result (295 GB
    # per GPU) exceeds HBM, so it is produced in time tiles into a reused buffer (tile = samples per tile)
    "cfg3": dict(fs=30000.0, n=18000000, channels=32, freq_limits=[1.7, 15000.0], vpo=10, output="power",
                 tile=2250000, desc="32ch/GPU x 30kHz x 10min, 128 scales, fp32 power, streamed in 8 time tiles"),
    # config 4 of BASELINE.json: one 24 h recording at 30 kHz (2.592e9 samples) time-sharded over 8 GPUs
    # = 3.24e8 samples per GPU; every step does the mean all-reduce, the NCCL halo exchange with both
    # neighbours and the tiled transform of the shard
    "cfg4": dict(fs=30000.0, n=324000000, channels=1, freq_limits=[1.7, 15000.0], vpo=10, output="power",
                 tile=36000000, time_shard=True, plan_n=2592000000,
                 desc="1ch x 30kHz, 3.24e8 samples per GPU (24 h over 8 GPUs), 128 scales, fp32 power, "
                      "time-sharded with NCCL halos, streamed in time tiles"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=None, help="override samples per channel")
    ap.add_argument("--channels", type=int, default=None, help="channels per GPU (default: workload's)")
    ap.add_argument("--tile", type=int, default=None, help="samples per time tile (tiled workloads)")
    ap.add_argument("--no-guard", action="store_true", help="plan without the execute-time accuracy guard")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def plan_frequencies(wl):
    from ghost_b200 import ContinuousWaveletTransform
    cwt = ContinuousWaveletTransform(dtype=np.float32)
    cwt.fs = wl["fs"]
    cwt.wavelet.fs = wl["fs"]
    return np.asarray(cwt.plan_frequencies(wl.get("plan_n", wl["n"]), freq_limits=wl["freq_limits"],
                                           voices_per_octave=wl["vpo"]))


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, through in-process NVML
    (an nvidia-smi subprocess per sample stalls the driver for hundreds of ms)."""

    def __init__(self, index=0, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz, self.power = None, []
        self.nv = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        if nv is None:
            return
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                if get_reasons is not None:
                    r = int(get_reasons(self.handle))
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
               "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
               "source": "nvml" if self.nv is not None else "unavailable"}
        if self.power:
            out["power_w_max"] = float(max(self.power))
        return out


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(wl, freqs, seconds_hint):
    """Time the oracle port (reference algorithm, scipy FFT, ThreadPool over scales) on a
    bounded sample: one channel, a prefix of the recording, all scales."""
    from multiprocessing import cpu_count
    from ghost_b200 import synth
    from oracle import cwt_oracle as orc
    n = wl["n"]
    # the reference needs roughly 0.3 us per coefficient per core; bound the sample
    est_full = 3.0e-7 * n * len(freqs) / max(1, min(cpu_count(), len(freqs)))
    n_s = n if est_full <= seconds_hint else max(int(n * seconds_hint / est_full), int(5 * 1.2 * 45000))
    n_s = min(n, n_s)
    x = synth.chirp_pink(n_s, wl["fs"], 0, np.float32)
    t0 = time.perf_counter()
    orc.cwt_amplitude(x, wl["fs"], frequencies=freqs, parallel=True)
    dt = time.perf_counter() - t0
    coeffs = n_s * len(freqs)
    return coeffs / dt, dt, n_s, cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = dict(WORKLOADS[args.workload])
    if args.samples:
        wl["n"] = int(args.samples)
    freqs = plan_frequencies(wl)
    from multiprocessing import cpu_count
    from ghost_b200 import synth
    from oracle import cwt_oracle as orc
    # per step: one channel x a bounded prefix x all scales, all host threads
    rate, dt, n_s, cores = cpu_sample(wl, freqs, 8.0)
    x = synth.chirp_pink(n_s, wl["fs"], 0, np.float32)
    for _ in range(max(0, args.warmup - 1)):
        orc.cwt_amplitude(x, wl["fs"], frequencies=freqs, parallel=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.cwt_amplitude(x, wl["fs"], frequencies=freqs, parallel=True)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_s * len(freqs) / dt
    sample = "1 channel x %d samples x %d scales per step (oracle port of the reference, parallel=True)" % (
        n_s, len(freqs))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ghost_b200 import Morse, synth, _lib
    from ghost_b200.engine import CwtPlan, scale_tables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")                  # keep NCCL's banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    wl = dict(WORKLOADS[args.workload])
    if args.channels:
        wl["channels"] = args.channels
    if args.samples:
        wl["n"] = args.samples
    fs, n, nch = wl["fs"], wl["n"], wl["channels"]
    freqs = plan_frequencies(wl)
    m = Morse(fs=fs)
    om = freqs / (fs / 2.0) * np.pi
    L = m.compute_lengths(om)
    k0, nt, terms = scale_tables(m, om, L)
    plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, output=wl["output"], device=local, guard=not args.no_guard)
    S = len(freqs)

    # synthetic channels of this rank (channel shard: rank r owns channels r*nch .. r*nch+nch-1);
    # distinct seeds for the first 8, then reuse (generation cost only)
    if n > 50000000:       # very long shards: tile a 2^24-sample block (generation cost only)
        blk = [synth.chirp_pink(1 << 24, fs, rank * nch + c, np.float32) for c in range(min(nch, 8))]
        base = [np.tile(b, n // b.size + 1)[:n] for b in blk]
    else:
        base = [synth.chirp_pink(n, fs, rank * nch + c, np.float32) for c in range(min(nch, 8))]
    x_host = torch.empty((nch, n), dtype=torch.float32).pin_memory()
    for c in range(nch):
        x_host[c] = torch.from_numpy(base[c % len(base)])
        if c >= len(base):
            x_host[c] = torch.roll(x_host[c], 1009 * c)
    x_dev = x_host.to(dev, non_blocking=True)
    tile = int(args.tile if args.tile else wl.get("tile", 0))
    out = plan.alloc_out(nch, tile if tile else n)
    step_means = plan.channel_means(x_dev) if tile else None
    torch.cuda.synchronize(dev)

    def one_pass():
        if wl.get("time_shard"):
            from ghost_b200 import sharding
            sharding.run_time_shard_tiled(plan, x_dev, rank, world, tile, out=out)
        elif tile:
            plan.execute_tiled(x_dev, tile, out=out, means=step_means)
        else:
            plan.execute(x_dev, out)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident timing --------------------------------------------------
    for _ in range(args.warmup):
        one_pass()
    barrier()
    plan.profile(True)
    plan.profile_read(reset=True)
    _lib.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    step_wall = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        one_pass()
        step_wall.append(time.perf_counter() - t_s)
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms_total = ev0.elapsed_time(ev1)
    if os.environ.get("GCWT_BENCH_VERBOSE"):
        print("rank %d host ms per step (launch side): %s; device total %.1f ms" % (
            rank, ["%.1f" % (t * 1e3) for t in step_wall], ms_total), file=sys.stderr, flush=True)
    launches = _lib.launch_count()
    gstats = plan.guard_stats()
    prof = plan.profile_read(reset=True)
    plan.profile(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    coeffs_rank = float(nch) * n * S
    value = coeffs_rank * world / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel family -------------------------------------------
    levels = plan.levels()
    out_el = 8 if wl["output"] == "complex" else 4
    interp_on = wl["output"] != "complex"
    # level 2 joins the interpolated classes when its bands allow the wide coarse grid (no direct launches then)
    min_interp_level = 2 if prof["fused_banded"][1] == 0 else 3
    n_interp = int((levels >= min_interp_level).sum()) if interp_on else 0
    n_banded = int((levels >= 0).sum()) - n_interp
    n_full = int((levels == -1).sum())
    fam_scales = {"fused_interp": n_interp, "fused_banded": n_banded, "fused_full": n_full}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    fams = {}
    for fam, ns in fam_scales.items():
        ms_f, ln_f = prof[fam]
        if ns == 0 or ms_f <= 0:
            continue
        # algorithmic bytes of this family's launches in one step: its output rows plus one read of
        # the (decimated) input per class; the input term is bounded by 4 B/sample per launch
        by = float(nch) * n * (ns * out_el + 4)
        fams[fam] = {"scales": ns, "launches_per_step": ln_f / args.steps, "ms_per_step": ms_f / args.steps,
                     "bytes_per_step": by, "achieved_gbs": by * args.steps / (ms_f * 1e-3) / 1e9,
                     "coeff_per_s": float(nch) * n * ns * args.steps / (ms_f * 1e-3)}
        fams[fam]["frac"] = fams[fam]["achieved_gbs"] / peak
    dom = max(fams, key=lambda k: fams[k]["ms_per_step"]) if fams else None
    whole_gbs = float(nch) * n * (S * out_el + 4) / (ms_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": (dom + "_kernel") if dom else None,
            "achieved": fams[dom]["achieved_gbs"] if dom else None, "peak": peak, "unit": "GB/s",
            "frac": fams[dom]["frac"] if dom else None, "traffic": None, "peak_source": peak_src,
            "share_of_step": (fams[dom]["ms_per_step"] / ms_step) if dom else None,
            "families": fams, "mean_pyramid_ms_per_step": prof["mean+pyramid"][0] / args.steps,
            "whole_step": {"achieved": whole_gbs, "frac": whole_gbs / peak}}
    # measured DRAM traffic (dram__bytes_read + dram__bytes_write of one ncu --set full capture of this
    # kernel family, profiles/traffic.json) scaled from the captured channel count to this run's
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file) and dom:
        try:
            tj = json.load(open(traffic_file))
            ratio = tj.get(dom + "_traffic_over_algorithmic")
            if ratio is not None:
                roof["traffic"] = ratio * fams[dom]["bytes_per_step"]
                roof["traffic_note"] = "per step over all %s launches; ncu ratio traffic/algorithmic = %.3f (%s)" % (
                    dom, ratio, tj.get("source", ""))
        except Exception:
            pass

    # ---- end to end through the host-buffer path ----------------------------------------
    e2e = None
    if not args.no_e2e and not tile:
        e2e = run_e2e(args, plan, x_host, dev, world, nch, n, S, out_el)

    # ---- CPU baseline (rank 0, N = 1) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, dt, n_s, cores = cpu_sample(wl, freqs, args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
               "sample": "1 channel x %d samples x %d scales (oracle port of the reference: scipy FFT overlap-add, "
                         "ThreadPool over scales)" % (n_s, S)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "channels_per_gpu": nch, "samples": n,
                       "scales": S, "fs": fs, "output": wl["output"],
                       "parallelism": ("time-shard x%d (NCCL halo exchange + mean all-reduce per step)" if wl.get("time_shard")
                                       else "channel-shard x%d") % world,
                       "l2_policy": "inputs (%.2f GB) and outputs (%.1f GB) per step exceed the 126 MB L2" % (
                           nch * n * 4 / 1e9, coeffs_rank * out_el / 1e9),
                       "scale_classes": {"band_limited_interpolated": n_interp, "band_limited": n_banded,
                                         "full_spectrum": n_full, "generic": int((levels == -2).sum())}},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "guard": {"enabled": not args.no_guard, "pairs_checked": gstats["checked"],
                      "pairs_recomputed_fp64": gstats["total"]},
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, plan, x_host, dev, world, nch, n, S, out_el):
    """Host buffers in, host buffers out: per step every channel group is copied H2D from
    pinned memory, transformed, and its coefficients copied D2H into a pinned ring."""
    import torch
    import torch.distributed as dist
    group = max(1, min(nch, int(2.0e9 // (n * S * out_el)) or 1))
    ring = [torch.empty((group, S, n), dtype=plan.torch_out_dtype).pin_memory() for _ in range(2)]
    dbuf = [plan.alloc_out(group, n) for _ in range(2)]
    xin = [torch.empty((group, n), dtype=torch.float32, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    done = [torch.cuda.Event() for _ in range(2)]
    computed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def one_step():
        i = 0
        for c0 in range(0, nch, group):
            g = min(group, nch - c0)
            b = i & 1
            main.wait_event(done[b])                       # ring slot free again
            xin[b][:g].copy_(x_host[c0:c0 + g], non_blocking=True)
            plan.execute(xin[b][:g], dbuf[b][:g])
            computed[b].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(computed[b])
                ring[b][:g].copy_(dbuf[b][:g], non_blocking=True)
                done[b].record(copy_stream)
            i += 1
        copy_stream.synchronize()

    steps = max(1, min(args.steps, 3))
    one_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / steps
    checksum = float(ring[0][0, 0, :1000].double().sum())
    return {"value": float(nch) * n * S * world / sec, "unit": UNIT, "h2d_bytes_per_step": int(nch * n * 4),
            "d2h_bytes_per_step": int(nch * n * S * out_el), "ms_per_step": sec * 1e3, "steps": steps,
            "channel_group": group, "api": "CwtPlan.execute on pinned host buffers, D2H of all coefficients",
            "checksum": checksum}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
