#!/usr/bin/env python
"""bench.py — headline benchmark of the Monte-Carlo decode path (BASELINE.json):
decoded information Gbit/s of the (133,171) K=7 code + FNN decoder (14-128-64-32-128 tanh),
2^24 frames per SNR point per GPU, encode -> BPSK+AWGN -> decode -> BER/FER count.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

A "step" is one pass of the whole hot path over 2^24 frames at one SNR point (2 dB).
  value  : whole-job info Gbit/s of the fused persistent tcgen05 pipeline (no per-frame HBM input:
           frames are a pure function of seed / SNR index / global frame index), CUDA events on the
           library's stream, max over ranks; the counters' exchange (NVLink peer memory, or ncclAllReduce) is
           issued by the library on the same stream behind the last launch and is inside the timed region.
           Timed region: W warm-up steps, barrier + synchronize, --preroll untimed launches (the clocks sag while
           the GPU idles in the barrier), one tiny exchange that aligns the ranks on the device, event, exactly K
           steps, the exchange, event, synchronize, barrier.
  e2e    : same metric through the host-buffer C-ABI call nvd_mc_run_host: info words start in
           pinned HOST memory, decisions + counters end in host memory, copies inside the region.
  --impl reference : the CPU oracle port of the reference path (numpy noise + torch-CPU fp32 MLP,
           all host threads) on a bounded sample per step; rank 0 only.
  extra legs (skipped with --no-legs): `sustained` (>= 3 s of back-to-back launches with clocks and power),
           `deep_ber` (configs[4], nvd_mc_sweep), `bitwise` + `mixed_snr_sweep` (configs[2]), `rnn_viterbi` (configs[3]),
           `seq_predict`, `decode_host` (the reference's own decode surface on host arrays).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

L = 7
FRAMES_PER_STEP = 1 << 24            # BASELINE.json configs[1]: 2^24 frames per SNR point on one B200
SNR_DB = 2.0                         # reference default (src/fnn_viterbi.py:699)
FLOP_PER_FRAME = 32256               # 2 * (14*128 + 128*64 + 64*32 + 32*128)   SURVEY.md §8d
METRIC = "decoded_info_gbit_per_s"
UNIT = "Gbit/s"
WORKLOAD = "configs[1]: (133,171) K=7 code, FNN decoder 14-128-64-32-128, 2^24 frames per SNR point per GPU, 2 dB"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.active, self.stop_flag = index, [], set(), False, False
        self.power = []
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self, force=False):
        """One NVML reading; kept when the sampler is active (or forced: the main thread samples the short headline
        region itself, while its launches run, so that the region never ends up without a reading)."""
        if not self.ok:
            return
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            if self.active or force:
                self.samples.append(float(mhz))
                try:
                    self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            self.sample()
            time.sleep(float(os.environ.get("NVD_BENCH_CLOCK_PERIOD_S", "0.005")))

    def reset(self):
        self.samples, self.power, self.reasons = [], [], set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power:
            out["power_w_median"] = float(np.median(self.power))
            out["power_w_max"] = float(np.max(self.power))
        return out


def cpu_pipeline_rate(model, frames_per_call: int, seconds_budget: float, threads=None):
    """Time the oracle port of the reference CPU path; returns (Gbit/s, frames_done, seconds, threads)."""
    import torch
    from oracle import encoders as oenc, pipeline as opipe
    if threads:
        torch.set_num_threads(threads)
    rng = np.random.default_rng(1)
    opipe.mc_step(model.Ws, model.bs, model.acts, L, oenc.CODE_133171, SNR_DB, 0, 1 << 16, noise="numpy", rng=rng)  # warm
    done, t0 = 0, time.perf_counter()
    while True:
        opipe.mc_step(model.Ws, model.bs, model.acts, L, oenc.CODE_133171, SNR_DB, done, frames_per_call, noise="numpy", rng=rng)
        done += frames_per_call
        dt = time.perf_counter() - t0
        if dt >= seconds_budget:
            break
    return done * L / dt / 1e9, done, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port — the
    reference itself is pure Python + TensorFlow 1.x and cannot run here)."""
    if rank != 0:
        return 0
    from nvd import weights
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm uses the whole host at every N
    torch.set_num_threads(os.cpu_count() or 1)
    model = weights.load_default_fnn()
    sample = 1 << 20                                   # bounded sample of the 2^24-frame step
    from oracle import encoders as oenc, pipeline as opipe
    rng = np.random.default_rng(1)
    for _ in range(max(args.warmup, 1)):
        opipe.mc_step(model.Ws, model.bs, model.acts, L, oenc.CODE_133171, SNR_DB, 0, sample, noise="numpy", rng=rng)
    t0 = time.perf_counter()
    for s in range(args.steps):
        opipe.mc_step(model.Ws, model.bs, model.acts, L, oenc.CODE_133171, SNR_DB, s * sample, sample, noise="numpy", rng=rng)
    dt = time.perf_counter() - t0
    value = args.steps * sample * L / dt / 1e9
    threads = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} frames per step (bounded sample of the 2^24-frame step)",
                   "noise": "numpy Gaussian (vectorised)", "mlp": "torch-CPU fp32 (MKL)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {sample} frames; oracle restatement of the reference CPU path "
                                   "(TensorFlow 1.x is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of one mlp_tc_kernel<MODE_FUSED> launch
# (profiles/r1_ncu_full_fused_summary.txt): the 40 KB weight image and the parameters, nothing per frame
NCU_DRAM_BYTES_PER_LAUNCH = 85632     # 85 760 and 85 504 B in the two captured launches (profiles/r2_ncu_full_fused_summary.txt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per step per GPU")
    ap.add_argument("--preroll", type=int, default=6,
                    help="untimed launches queued between the barrier + synchronize and the first timed step (the GPU's clocks "
                         "sag while it idles in the barrier)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-staged", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="headline + e2e only (skip sustained / deep-BER / bitwise / rnn / seq legs)")
    ap.add_argument("--deep-min-errors", type=int, default=100, help="deep-BER stop rule: frame errors per SNR point")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from nvd import dist
    rank, world, local = dist.env_rank_world()
    if args.impl == "reference":
        return run_reference(args, rank, world)

    from nvd import montecarlo, weights
    if world > 1:
        dist.init("nccl")
    model = weights.load_default_fnn()
    sim = montecarlo.Simulator(device=local if world > 1 else 0)
    sim.comm_init(rank, world)                          # the library's own NCCL communicator (nvd_comm_init)
    h = sim.load_mlp(model)
    frames = args.frames
    frame0 = rank * frames                              # weak scaling: every rank owns its own 2^24-frame range per step
    info = sim.device_info()
    peaks = measured_peaks()

    sampler = ClockSampler(local if world > 1 else 0)
    sampler.start()

    # ---------------- device-resident throughput: fused persistent kernel ---------------------------
    # everything the timed region needs exists BEFORE the warm-up, so that the GPU idles only for the barrier +
    # synchronize the contract asks for (after ~5 ms of idling the first four launches run 1-9 % slow: tools/diag/gap_probe.py)
    counters = sim.zeros(32)
    align = sim.zeros(32)
    e0, e1, e2 = sim.event(), sim.event(), sim.event()
    from nvd import _lib as _lb
    for w in range(args.warmup):
        sim.run_async(h, "133171", L, [SNR_DB], frames, counters, frame0 + w * world * frames, mode="fused")
    sim.allreduce_counters(counters, 4)                 # warm the communicator's first-use set-up
    sim.sync()
    dist.barrier()
    sim.sync()
    sampler.active = True
    _lb.check(sim.lib.nvd_memset(sim.ctx, counters.ptr, 0, 32))       # asynchronous, on the library's stream
    # The barrier + synchronize above idle the GPU, and after an idle gap the first four launches run 1-12 % slow while
    # the clocks come back (tools/diag/gap_probe.py: 1.065 / 0.988 / 0.973 / 0.962 ms against 0.955).  A Monte-Carlo
    # job never idles between its launches, so the timed K steps start behind `preroll` more untimed warm-up launches
    # queued on the same stream (their counters are discarded).
    for w in range(args.preroll):
        sim.run_async(h, "133171", L, [SNR_DB], frames, align, frame0 + w * world * frames, mode="fused")
    # Ranks leave the host barrier tens of microseconds apart and run the pre-roll at slightly different speeds; one
    # (tiny) exchange on the library's stream right in front of the first event makes every rank's stream pass the
    # same point together, so the regions start aligned on the DEVICE (the gap it leaves is ~20 us: no clock sag).
    sim.allreduce_counters(align, 4)
    sim.record(e0)
    for s in range(args.steps):
        sim.run_async(h, "133171", L, [SNR_DB], frames, counters, frame0 + s * world * frames, mode="fused")
    sim.record(e1)
    # the path's one exchange step: sum of uint64[4] over all ranks (NVLink peer memory, or ncclAllReduce) on the library's stream, directly on the device
    # counters, queued behind the last launch (no host bounce) — inside the timed region
    sim.allreduce_counters(counters, 4)
    sim.record(e2)
    # everything above is queued, nothing has been waited for: read the clocks while the K steps run
    t_region = (args.preroll + args.steps) * 0.9e-3
    for i in range(4):
        time.sleep(t_region * (0.3 if i == 0 else 0.15))
        sampler.sample(force=True)
    sim.sync()
    fused_ms = sim.elapsed_ms(e0, e2)
    ar_ms = sim.elapsed_ms(e1, e2) if world > 1 else 0.0
    kernels_ms = sim.elapsed_ms(e0, e1)
    total_counts = counters.download(np.uint64, (1, 4))
    sampler.active = False
    dist.barrier()
    step_ms = dist.all_reduce_max(fused_ms / args.steps)
    kernel_ms = dist.all_reduce_max(kernels_ms / args.steps)
    kernel_ms_min = -dist.all_reduce_max(-kernels_ms / args.steps)        # fastest rank: the spread is GPU-to-GPU variation
    ar_ms_min = -dist.all_reduce_max(-ar_ms)                              # the exchange as seen by the LAST rank to arrive
    value = world * frames * L / (step_ms * 1e-3) / 1e9

    # ---------------- end to end: host buffers through nvd_mc_run_host -------------------------------
    e2e_value = e2e_ms = e2e_co_ms = None
    c_e2e = []
    if not args.no_e2e:
        info_host = sim.pinned(frames)
        pred_host = sim.pinned(frames)
        table = montecarlo.class_to_packed(np.arange(1 << L), L).astype(np.uint8)   # class -> time-packed info word
        info_host[:] = table[(frame0 + np.arange(frames, dtype=np.int64)) & ((1 << L) - 1)]
        e2e_steps = max(3, min(args.steps, 20))
        # the call a user makes: nvd_mc_run_host through ctypes (arguments prepared once, Simulator.host_step)
        step_pred = sim.host_step(h, "133171", L, SNR_DB, info_host, pred_host, frame0=frame0, precision="bf16", mode="fused")
        step_cnt = sim.host_step(h, "133171", L, SNR_DB, info_host, None, frame0=frame0, precision="bf16", mode="fused")
        for _ in range(3):
            step_pred()
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            c_e2e = step_pred()
        e2e_ms = dist.all_reduce_max((time.perf_counter() - t0) * 1e3 / e2e_steps)
        e2e_value = world * frames * L / (e2e_ms * 1e-3) / 1e9
        c_e2e = c_e2e.copy()
        # the same call without the per-frame decisions coming back (a Monte-Carlo caller only wants the counters)
        step_cnt()
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            step_cnt()
        e2e_co_ms = dist.all_reduce_max((time.perf_counter() - t0) * 1e3 / e2e_steps)

    # ---------------- staged pipeline: per-kernel HBM numbers (explains the fused number) ---------------
    staged = None
    if not args.no_staged:
        sim.set_stage_timing(True)
        c2 = sim.zeros(32)
        for _ in range(2):
            sim.run_async(h, "133171", L, [SNR_DB], frames, c2, frame0, mode="staged", precision="bf16")
        ms = sim.stage_ms()
        sim.set_stage_timing(False)
        alg_bytes = {"source": 1, "encode": 3, "channel": 30, "decode": 29, "count": 2}   # per frame, SURVEY.md §8d
        staged = {k: {"ms": ms[k], "GB/s": alg_bytes[k] * frames / (ms[k] * 1e-3) / 1e9 if ms[k] > 0 else None,
                      "frac_of_hbm": (alg_bytes[k] * frames / (ms[k] * 1e-3) / 1e9 / peaks["hbm_gbs"]) if ms[k] > 0 else None}
                  for k in ms}
        staged["total_ms"] = sum(ms.values())
        staged["gbit_per_s"] = frames * L / (staged["total_ms"] * 1e-3) / 1e9

    clocks = sampler.summary()             # clocks of the headline timed region only

    # ---------------- sustained: >= 3 s of back-to-back fused launches (clock / power state over seconds) -----------
    sustained = None
    if not args.no_legs:
        n_sus = max(int(3200.0 / max(step_ms, 1e-3)), 100)
        c_sus = sim.zeros(32)
        sampler.reset()
        dist.barrier()
        sampler.active = True
        sim.record(e0)
        for s_ in range(n_sus):
            sim.run_async(h, "133171", L, [SNR_DB], frames, c_sus, frame0 + (args.steps + s_) * world * frames, mode="fused")
        sim.allreduce_counters(c_sus, 4)
        sim.record(e1)
        sim.sync()
        sampler.active = False
        sus_ms = dist.all_reduce_max(sim.elapsed_ms(e0, e1))
        sus_tflops = FLOP_PER_FRAME * frames * n_sus / (sus_ms * 1e-3) / 1e12
        sustained = {"launches": n_sus, "seconds": sus_ms * 1e-3, "value": world * frames * n_sus * L / (sus_ms * 1e-3) / 1e9, "unit": UNIT,
                     "ms_per_step": sus_ms / n_sus, "tflops_per_gpu": sus_tflops,
                     "frac_of_bf16_sustained_peak": sus_tflops / peaks["bf16_sustained"], "clocks": sampler.summary()}

    # ---------------- BASELINE.json configs[4]: deep-BER run with the stop rule, one C-ABI call (nvd_mc_sweep) ------
    deep = None
    if not args.no_legs:
        deep_db = [5.0, 6.0, 7.0, 8.0]
        dist.barrier()
        t0 = time.perf_counter()
        dc = sim.sweep(h, "133171", L, deep_db, min_frame_errors=args.deep_min_errors, max_frames=1 << 34,
                       chunk_frames=world << 26, mode="fused", world=world)
        deep_s = dist.all_reduce_max(time.perf_counter() - t0)
        pts = []
        for db_, c_ in zip(deep_db, dc):
            lo_, hi_ = montecarlo.wilson(float(c_[0]), float(c_[3]))
            pts.append({"snr_db": db_, "frames": int(c_[2]), "bit_errors": int(c_[0]), "frame_errors": int(c_[1]),
                        "ber": float(c_[0]) / max(float(c_[3]), 1.0), "fer": float(c_[1]) / max(float(c_[2]), 1.0), "ber_ci95": [lo_, hi_]})
        tot_frames = int(dc[:, 2].sum())
        deep = {"workload": "configs[4]: (133,171) FNN deep-BER, stop at >= %d frame errors per point or 2^34 frames, "
                            "chunks of 2^26 frames per GPU, counters all-reduced in the library after every chunk" % args.deep_min_errors,
                "points": pts, "frames_total": tot_frames, "seconds": deep_s, "gbit_per_s": tot_frames * L / deep_s / 1e9,
                "frac_of_headline": tot_frames * L / deep_s / 1e9 / value}

    # ---------------- BASELINE.json configs[2]: per-bit predictor + mixed-SNR decoder, 0..8 dB, frames sharded over ranks --
    sweep_db = [float(d) for d in range(0, 9)]
    bitwise = None
    if not args.no_legs:
        from nvd import weights as _wb
        mb = _wb.load_default_bitwise()
        hb = sim.load_mlp(mb)
        fb = 1 << 21                                       # frames per SNR point per GPU
        cb_ = sim.zeros(32 * len(sweep_db))
        sim.run_bitwise_async(hb, 512, [2.0], 1 << 16, cb_, 0, mode="fused")       # warm
        sim.sync()
        cb_ = sim.zeros(32 * len(sweep_db))
        dist.barrier()
        sim.record(e0)
        sim.run_bitwise_async(hb, 512, sweep_db, fb, cb_, rank * fb, mode="fused", trunc=True)
        sim.allreduce_counters(cb_, 4 * len(sweep_db))
        sim.record(e1)
        sim.sync()
        bw_ms = dist.all_reduce_max(sim.elapsed_ms(e0, e1))
        cbw = cb_.download(np.uint64, (len(sweep_db), 4))
        n_fr = world * fb * len(sweep_db)
        bitwise = {"workload": "configs[2]: fnn_viterbi_bitwise 1036-30-2 (ReLU) predictor of bit 0, L=512 + 6 state bits, "
                               "int32-truncated channel (reference data sets), 0..8 dB, fused K-streamed tcgen05 pipeline",
                   "frames_per_point_per_gpu": fb, "ms": bw_ms, "mframes_per_s": n_fr / (bw_ms * 1e-3) / 1e6,
                   "decoded_mbit_per_s": n_fr / (bw_ms * 1e-3) / 1e6, "gnormals_per_s": n_fr * 1036 / (bw_ms * 1e-3) / 1e9,
                   "tflops": n_fr * 62280 / (bw_ms * 1e-3) / 1e12,
                   "bit_error_rate": [float(c_[0]) / max(float(c_[3]), 1.0) for c_ in cbw], "weights": mb.meta.get("path", "seed-fixed init")}
        if world == 1:                                     # staged: per-stage HBM rates (algorithmic bytes)
            sim.set_stage_timing(True)
            cs_ = sim.zeros(32)
            fs_ = 1 << 18
            for _ in range(2):
                sim.run_bitwise_async(hb, 512, [2.0], fs_, cs_, 0, mode="staged", trunc=True)
            msb = sim.stage_ms()
            sim.set_stage_timing(False)
            algb = {"source": 64 + 64 + 65, "encode": 65 + 130, "channel": 130 + 2072, "decode": 2072 + 1, "count": 1 + 1}
            bitwise["staged"] = {k: {"ms": msb[k], "GB/s": algb[k] * fs_ / (msb[k] * 1e-3) / 1e9 if msb[k] > 0 else None,
                                     "frac_of_hbm": algb[k] * fs_ / (msb[k] * 1e-3) / 1e9 / peaks["hbm_gbs"] if msb[k] > 0 else None}
                                 for k in msb}
            bitwise["staged"]["frames"] = fs_
            bitwise["staged"]["mframes_per_s"] = fs_ / (sum(msb.values()) * 1e-3) / 1e6

    mixed = None
    if not args.no_legs:
        cm_ = sim.zeros(32 * len(sweep_db))
        dist.barrier()
        sim.record(e0)
        sim.run_async(h, "133171", L, sweep_db, frames, cm_, frame0, mode="fused")
        sim.allreduce_counters(cm_, 4 * len(sweep_db))
        sim.record(e1)
        sim.sync()
        mx_ms = dist.all_reduce_max(sim.elapsed_ms(e0, e1))
        cmx = cm_.download(np.uint64, (len(sweep_db), 4))
        mixed = {"workload": "configs[2]: fnn_viterbi_mixed_snr decoder (one model trained over -2..9 dB) tested at 0..8 dB, "
                             "2^24 frames per point per GPU, one all-reduce for the sweep",
                 "ms": mx_ms, "gbit_per_s": world * frames * len(sweep_db) * L / (mx_ms * 1e-3) / 1e9,
                 "ber": [float(c_[0]) / max(float(c_[3]), 1.0) for c_ in cmx], "fer": [float(c_[1]) / max(float(c_[2]), 1.0) for c_ in cmx]}

    # ---------------- BASELINE.json configs[3]: rnn_viterbi as a Monte-Carlo pipeline (nvd_mc_run_lstm) ---------------
    other = None
    if not args.no_legs:
        from nvd import weights as _w
        lm = _w.load_default_lstm()
        hl = sim.load_lstm(lm)
        Ll, Bl = 100, info["sm_count"] * 5 * 128 * 4
        other = {"workload": "configs[3]: rnn_viterbi LSTM(20) seq2seq on (133,171) L=100 blocks: Philox info -> K1 -> K2 (int32-"
                             "truncated) -> K4 (200 + 200 steps) -> K5, frames sharded over ranks, 2 dB", "frames_per_gpu": Bl}
        for name, fm, teacher, reps in (("tcgen05_teacher_forced", 2, True, 5), ("tcgen05_autoregressive", 2, False, 5), ("fp32_pipe_teacher_forced", 1, True, 1)):
            cl_ = sim.zeros(32)
            sim.run_lstm_async(hl, Ll, [2.0], Bl, cl_, rank * Bl, fast_math=fm, teacher_forced=teacher)
            sim.sync()
            cl_ = sim.zeros(32)
            dist.barrier()
            sim.record(e0)
            for r_ in range(reps):
                sim.run_lstm_async(hl, Ll, [2.0], Bl, cl_, (r_ * world + rank) * Bl, fast_math=fm, teacher_forced=teacher)
            sim.allreduce_counters(cl_, 4)
            sim.record(e1)
            sim.sync()
            ms_l = dist.all_reduce_max(sim.elapsed_ms(e0, e1)) / reps
            cl = cl_.download(np.uint64, (4,))
            other[name] = {"ms": ms_l, "gbit_per_s": world * Bl * Ll / (ms_l * 1e-3) / 1e9, "mframes_per_s": world * Bl / (ms_l * 1e-3) / 1e6,
                           "ber": float(cl[0]) / max(float(cl[3]), 1.0)}
        if world == 1:
            sim.set_stage_timing(True)
            cl_ = sim.zeros(32)
            sim.run_lstm_async(hl, Ll, [2.0], Bl, cl_, 0, fast_math=2, teacher_forced=True)
            other["stage_ms"] = sim.stage_ms()
            sim.set_stage_timing(False)

    # ---------------- the reference's own decode surface: predictions.eval(feed_dict={x: test_set}) on HOST arrays ------
    decode_host = None
    if not args.no_legs and not args.no_e2e:
        nd = 1 << 22
        y_host = sim.pinned(nd * 14 * 4).view(np.float32).reshape(nd, 14)
        rng_d = np.random.default_rng(7)
        y_host[:] = rng_d.standard_normal((nd, 14), dtype=np.float32)
        hard_host = sim.pinned(nd * 4).view(np.int32)
        from nvd import _lib as _ld
        def dec_step(prec):
            _ld.check(sim.lib.nvd_mlp_decode_host(sim.ctx, h, y_host.ctypes.data, nd, _ld.PRECISIONS[prec], None, hard_host.ctypes.data))
        dec_step("bf16")
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dec_step("bf16")
        dh_ms = dist.all_reduce_max((time.perf_counter() - t0) * 1e3 / 3)
        decode_host = {"api": "nvd_mlp_decode_host (fnn_viterbi.predict): fp32 rows in pinned host memory -> int32 decisions in host memory; "
                              "double-buffered H2D / convert + tcgen05 forward / D2H on three streams",
                       "frames": nd, "ms": dh_ms, "value": world * nd * L / (dh_ms * 1e-3) / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": nd * 56, "d2h_bytes_per_step": nd * 4, "h2d_gb_per_s": nd * 56 / (dh_ms * 1e-3) / 1e9}

    # ---------------- sliding-window decoder (fnn_viterbi_seq_predict, 200-500x4-2) on the wide-layer tcgen05 kernel ---
    seq = None
    if not args.no_legs and world == 1:
        from nvd import weights as _w2, _lib as _l2
        dims_s = [200, 500, 500, 500, 500, 2]
        rng = np.random.default_rng(3)
        Ws_s = [(rng.standard_normal((a, b)) * np.sqrt(2.0 / a)).astype(np.float32) for a, b in zip(dims_s[:-1], dims_s[1:])]
        bs_s = [np.zeros(b, np.float32) for b in dims_s[1:]]
        hs = sim.load_mlp(_w2.MlpModel(Ws_s, bs_s, [1] * 5))
        n_win = 1 << 19
        n_sym = 2 * (n_win - 1) + 200
        d_stream = sim.to_device(rng.standard_normal(n_sym).astype(np.float32))
        d_bits = sim.alloc(n_win)
        def seq_step():
            _l2.check(sim.lib.nvd_window_decode(sim.ctx, hs, d_stream.ptr, _l2.DTYPES["f32"], n_sym, 2, n_win,
                                                _l2.PRECISIONS["bf16"], None, d_bits.ptr, 1))
        seq_step(); sim.sync()
        e0, e1 = sim.event(), sim.event()
        sim.record(e0)
        for _ in range(3):
            seq_step()
        sim.record(e1); sim.sync()
        ms_s = sim.elapsed_ms(e0, e1) / 3
        mac = sum(a * b for a, b in zip(dims_s[:-1], dims_s[1:]))
        seq = {"workload": "fnn_viterbi_seq_predict: 200-500x4-2 tanh MLP over stride-2 windows, nvd_window_decode bf16", "windows": n_win,
               "ms": ms_s, "mbit_per_s": n_win / (ms_s * 1e-3) / 1e6, "tflops": 2 * mac * n_win / (ms_s * 1e-3) / 1e12,
               "frac_of_bf16_peak": 2 * mac * n_win / (ms_s * 1e-3) / 1e12 / peaks["bf16_burst"]}

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank != 0:
        sim.close()
        dist.shutdown()
        return 0

    # ---------------- roofline of the dominant kernel (the fused tcgen05 pipeline kernel) ----------------
    achieved_tflops = FLOP_PER_FRAME * frames / (kernel_ms * 1e-3) / 1e12
    long_region = fused_ms > 1000.0
    peak = peaks["bf16_sustained"] if long_region else peaks["bf16_burst"]
    roofline = {"bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tflops / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "kernel": "mlp_tc_kernel<MODE_FUSED>",
                "peak_source": peaks["source"] + (", sustained" if long_region else ", burst"),
                "note": "traffic = dram bytes per launch from the committed ncu --set full capture (weights only: the fused "
                        "kernel generates its frames on chip); not HBM-bound.  Bound by the instruction stream of its 16 "
                        "epilogue warps per SM (~1 160 instructions per frame, 168 of them MUFU: 140 tanh after the FMA-pipe "
                        "polynomial takes 3/8 of them + 28 for the channel noise); MUFU pipe 65 %, issue slots 56 %, tensor "
                        "pipe 30 % busy (DESIGN.md §4)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        v, done, dt, threads = cpu_pipeline_rate(model, 1 << 20, 12.0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{done} frames in {dt:.1f} s (oracle port: numpy noise + torch-CPU fp32 MLP + popcount)"}

    ber = float(total_counts[0, 0]) / max(float(total_counts[0, 3]), 1.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": frames, "snr_db": SNR_DB, "mode": "fused",
                   "weights": model.meta.get("path", "seed-fixed init"), "seed": sim.seed,
                   "l2": "fused kernel has no per-frame HBM input (40 KB of weights); staged buffers are 1 GB >> 126 MB L2",
                   "sm_count": info["sm_count"],
                   "preroll": f"{args.preroll} untimed launches between the barrier + synchronize and the first timed step "
                              "(clock recovery after the idle gap, tools/diag/gap_probe.py); the K timed steps follow back to back"},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(frames), "d2h_bytes_per_step": int(frames) + 32,
                "ms_per_step": e2e_ms,
                "counters_only": None if e2e_co_ms is None else {"value": world * frames * L / (e2e_co_ms * 1e-3) / 1e9, "ms_per_step": e2e_co_ms,
                                                                 "d2h_bytes_per_step": 32, "note": "same call with pred_host = NULL"},
                "api": "nvd_mc_run_host, NVD_MC_FUSED, pinned host buffers: ONE launch of the fused tcgen05 kernel; info words arrive by sliced H2D copies while it runs (flag per slice), decisions leave by D2H copies queued as the kernel reports slices complete; counters D2H (sliced launches instead under a profiler)"},
        # the K fused-kernel launches of the timed region, and the two small kernels of the peer-memory exchange behind them
        "gpu_launches": args.steps + (2 if (world > 1 and sim.p2p) else 0),
        "clocks": clocks,
        "ber": ber, "frame_errors": int(total_counts[0, 1]), "frames_total": int(total_counts[0, 2]),
        "allreduce_ms": ar_ms,
        "ranks": {"kernel_ms_per_step_max": kernel_ms, "kernel_ms_per_step_min": kernel_ms_min,
                  "exchange_ms_rank0": ar_ms, "exchange_ms_min_over_ranks": ar_ms_min,
                  "note": "kernel = e0..e1 (K launches) per rank; exchange = e1..e2: the waiting for slower ranks is inside it on "
                          "every rank but the slowest, whose figure is the cost of the exchange itself"},
        "comm": sim.comm_info(),
        "staged": staged,
        "sustained": sustained,
        "deep_ber": deep,
        "bitwise": bitwise,
        "mixed_snr_sweep": mixed,
        "rnn_viterbi": other,
        "seq_predict": seq,
        "decode_host": decode_host,
        "e2e_counters": [int(x) for x in c_e2e],
    }
    print(json.dumps(line), flush=True)
    sim.close()
    dist.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
