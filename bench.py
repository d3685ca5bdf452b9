#!/usr/bin/env python
"""bench.py -- headline benchmark of the receive-chain hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): IQBaseBand<float>(64 taps, shift 100 kHz, 20 MS/s ->
48 kHz) + FMDemod on complex-float IQ; one "step" = one pass of the fused chain over a batch of
256 buffers of 2^20 samples (256 Mi samples, 2 GiB > L2, so every step streams from HBM).
Metric: input IQ Msamples/s.

  value    : device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e      : the same through the host-pointer C-ABI call (sdrg_rxchain_process) from pinned host
             memory, H2D of the step's input and D2H of its audio inside the timed region
  roofline : the dominant kernel (iqbb accumulate) timed with CUDA events on its launching stream
  cpu_baseline : the oracle port (the reference has no float instantiation) on one host core

N > 1: one process per GPU, every rank runs its own independent stream (weak scaling, no data-path
collective); the demodulated audio of all ranks is gathered with NCCL (all_gather batched over 8 steps, overlapped with the next steps).
"""
import argparse
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

from libsdr_b200 import synth  # noqa: E402

METRIC = "iq_msamples_per_s_iqbaseband_fmdemod"
UNIT = "Msamples/s"


def workload():
    c = dict(synth.C2)
    c["n_buffers"] = 256        # one step = 256 buffers of 2^20 samples = 2 GiB of cf32 per GPU
    return c


def config_block(c, n_gpus, extra=None):
    cfg = {"workload": "c2: IQBaseBand<float> 64-tap FIR, shift 100 kHz, 20 MS/s -> 48 kHz (ss=416) + FMDemod",
           "scalar": "cf32", "order": c["order"], "sample_rate": c["Fs"], "output_rate": c["oFs"],
           "buffer_size": c["buffer_size"], "buffers_per_step": c["n_buffers"],
           "samples_per_step_per_gpu": c["buffer_size"] * c["n_buffers"],
           "l2_policy": "inputs larger than L2 (%d MiB per step per GPU)" % (c["buffer_size"] * c["n_buffers"] * 8 >> 20),
           "input": "3 tones + uniform noise (synth.c2_input), a 4 Mi-sample segment tiled to the batch",
           "parallelism": "independent streams per GPU (replicas); NCCL all_gather of the audio, 8 steps per collective, overlapped" if n_gpus > 1 else "single GPU"}
    if extra:
        cfg.update(extra)
    return cfg


def make_input(c):
    seg = synth.c2_input(4 << 20)
    reps = (c["buffer_size"] * c["n_buffers"]) // seg.shape[0]
    return np.tile(seg, (reps, 1))


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---- CPU baselines ---------------------------------------------------------------------------------
def cpu_port_run(c, x, threads, n_buffers):
    """The oracle port of the float chain on `threads` independent streams; returns Msamples/s."""
    from oracle import oracle as orc
    bs = c["buffer_size"]

    def one(res, k):
        o = orc.IQBaseBand(orc.F32, c["Fc"], c["Ff"], c["width"], c["order"], c["sub_sample"], c["oFs"])
        o.config(c["Fs"], bs)
        fm = orc.FMDemod(orc.F32)
        n = 0
        for b in range(n_buffers):
            y = o.process(x[(b % c["n_buffers"]) * bs:((b % c["n_buffers"]) + 1) * bs])
            if y.shape[0]:
                fm.process(y, inplace=True)
            n += bs
        res[k] = n

    res = [0] * threads
    ths = [threading.Thread(target=one, args=(res, k)) for k in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(res) / dt / 1e6, dt


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference has no float IQBaseBand/FMDemod (they do not compile/link), so the float workload is
    timed on the oracle port, one independent stream per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = workload()
    x = make_input(c)
    cores = os.cpu_count() or 1
    nb = 2                                           # buffers per thread per step (bounded sample)
    for _ in range(max(args.warmup, 0)):
        cpu_port_run(c, x, cores, 1)
    vals, tot = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_port_run(c, x, cores, nb)
        vals.append(v); tot += dt
    value = float(np.mean(vals))
    sample = "%d threads x %d buffers of %d samples per step" % (cores, nb, c["buffer_size"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (computed in f64 on CPU)",
            "data": "synthetic", "config": config_block(c, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _time_steps(fn, steps, warmup, barrier):
    import torch
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def secondary_workloads(world, rank, dev, dist, args):
    """Short measurements of the other BASELINE configs (not the headline line): the sharded
    2048-channel bank (C5) on all ranks, and -- single GPU only -- the int16 chain (C1) and the FFT filter (C3)."""
    import torch
    from libsdr_b200 import parallel
    from libsdr_b200.nodes import ChannelBank, IQBaseBand, RxChain, FilterNode, DEMOD_FM
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    try:   # C5: 2048 channels sharded by contiguous channel ranges, outputs gathered with NCCL
        c = dict(synth.C5)
        bs, nb = c["buffer_size"], 2
        fc_all = synth.bank_frequencies(c["channels"], c["Fs"])
        fc, lo, hi = parallel.bank_frequencies_for_rank(rank, world, fc_all)
        x = torch.from_numpy(synth.bank_input(bs, dict(c, channels=16))).to(dev).repeat(nb, 1)
        bank = ChannelBank("s16", fc, None, c["width"], c["order"], c["sub_sample"], c["oFs"])
        bank.config(sample_rate=c["Fs"], buffer_size=bs)
        n_out = bank.outputs_for(bs * nb) + 1
        bufs = {"fm": torch.zeros((hi - lo, n_out), dtype=torch.int16, device=dev),
                "am": torch.zeros((hi - lo, n_out), dtype=torch.int16, device=dev)}

        def step():
            r = bank.process(x, bs, want=("fm", "am"), out=bufs)
            if world > 1:
                parallel.gather_channel_outputs(bufs["fm"], c["channels"])
                parallel.gather_channel_outputs(bufs["am"], c["channels"])
            return r

        ms = maxms(_time_steps(step, 3, 2, barrier))
        imad_peak = 148 * 64 * 1.965e9           # FMA-heavy pipe: 64 IMAD lanes / clk / SM
        out["c5_bank"] = {"workload": "2048-channel IQBaseBand<int16> bank (15 taps, 100 MS/s -> 48 kHz) + FM + AM, channels "
                                      "sharded over %d GPU(s), NCCL all_gather of the audio" % world,
                          "input_msamples_per_s": bs * nb / ms / 1e3, "channel_msamples_per_s": c["channels"] * bs * nb / ms / 1e3,
                          "ms_per_step": ms, "buffers_per_step": nb, "scaling": "strong (channels fixed)",
                          "bound": "FMA-heavy (IMAD) pipe", "imad_pipe_frac_est": c["channels"] * bs * nb / (ms * 1e-3) * (3 * 14 + 4) / (imad_peak * world)}
    except Exception as e:  # pragma: no cover
        out["c5_bank"] = {"error": str(e)[:200]}
    if world > 1 or args.no_secondary:
        return out
    try:   # C1: int16, 15 taps, 2.4 MS/s -> 48 kHz + FM
        c = dict(synth.C1)
        bs, nb = c["buffer_size"], 2048
        x = torch.from_numpy(synth.c1_input(16 * bs)).to(dev).repeat(nb // 16, 1)
        bb = IQBaseBand("s16", c["Fc"], c["Ff"], c["width"], c["order"], c["sub_sample"], c["oFs"])
        bb.config(sample_rate=c["Fs"], buffer_size=bs)
        ch = RxChain(bb, DEMOD_FM)
        ms = _time_steps(lambda: ch.process(x, bs), 5, 2, barrier)
        out["c1_int16"] = {"workload": "IQBaseBand<int16> 15 taps, 2.4 MS/s -> 48 kHz + FMDemod, %d buffers of %d" % (nb, bs),
                           "msamples_per_s": bs * nb / ms / 1e3, "ms_per_step": ms,
                           "algorithmic_gbs": (4 + 2 / 50) * bs * nb / ms / 1e6, "bound": "FMA-heavy (IMAD) pipe (bit-exact path)",
                           "imad_pipe_frac_est": bs * nb / (ms * 1e-3) * (3 * 14 + 4) / (148 * 64 * 1.965e9)}
    except Exception as e:  # pragma: no cover
        out["c1_int16"] = {"error": str(e)[:200]}
    try:   # SURVEY 8(f) rows next to the path: RTL-style cu8 input with AutoCast fused into the load, and the real-input BaseBand<int16>
        from libsdr_b200 import _lib as L
        from libsdr_b200.nodes import BaseBand
        c = dict(synth.C1)
        bs, nb = c["buffer_size"], 2048
        g = torch.Generator(device="cpu"); g.manual_seed(0x5D12)
        x8 = torch.randint(0, 256, (nb * bs, 2), dtype=torch.uint8, generator=g).to(dev)
        bb = IQBaseBand("s16", c["Fc"], c["Ff"], c["width"], c["order"], c["sub_sample"], c["oFs"])
        bb.setInputType(L.T_CU8); bb.config(sample_rate=c["Fs"], buffer_size=bs)
        ch = RxChain(bb, DEMOD_FM)
        ms = _time_steps(lambda: ch.process(x8, bs), 5, 2, barrier)
        out["c1_cu8_fused_autocast"] = {"workload": "complex uint8 -> AutoCast fused into IQBaseBand<int16> (15 taps, ss=50) + FMDemod, %d buffers of %d" % (nb, bs),
                                        "msamples_per_s": bs * nb / ms / 1e3, "ms_per_step": ms,
                                        "algorithmic_gbs": (2 + 2 / 50) * bs * nb / ms / 1e6, "bound": "FMA-heavy (IMAD) pipe (bit-exact path)"}
        xr = torch.randint(-32768, 32768, (nb * bs,), dtype=torch.int16, generator=g).to(dev)
        rb = BaseBand(300e3, 300e3, 50e3, 32, 50); rb.config(sample_rate=c["Fs"], buffer_size=bs)
        chr_ = RxChain(rb, DEMOD_FM)
        ms = _time_steps(lambda: chr_.process(xr, bs), 5, 2, barrier)
        out["real_baseband_int16"] = {"workload": "BaseBand<int16> on a real int16 stream (32 taps, ss=50) + FMDemod, %d buffers of %d" % (nb, bs),
                                      "msamples_per_s": bs * nb / ms / 1e3, "ms_per_step": ms,
                                      "algorithmic_gbs": (2 + 2 / 50) * bs * nb / ms / 1e6, "bound": "FMA-heavy (IMAD) pipe (bit-exact path)"}
    except Exception as e:  # pragma: no cover
        out["next_rows"] = {"error": str(e)[:200]}
    try:   # C3: FFT-convolution filter, block 4096
        c = dict(synth.C3)
        nb = c["n_buffers"]
        x = torch.from_numpy(synth.c2_input(c["buffer_size"])).to(dev).repeat(nb, 1).view(torch.complex64).reshape(-1)
        f = FilterNode(c["block"]); f.addFilter(c["fmin"], c["fmax"]); f.config(sample_rate=c["Fs"], buffer_size=c["block"])
        ms = _time_steps(lambda: f.process(x), 5, 2, barrier)
        n = x.shape[0]
        out["c3_filter"] = {"workload": "FilterNode<float> block 4096 (FFT 8192), 1 band-pass filter, %d buffers of 2^20" % nb,
                            "msamples_per_s": n / ms / 1e3, "ms_per_step": ms, "algorithmic_gbs": 16 * n / ms / 1e6,
                            "bound": "shared memory / FP32 (Stockham FFT)"}
    except Exception as e:  # pragma: no cover
        out["c3_filter"] = {"error": str(e)[:200]}
    return out


# ---- our arm -----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C1/C3 side measurements")
    ap.add_argument("--buffers", type=int, default=0, help="buffers per step (default: the workload's n_buffers)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import ctypes as C
    import torch
    from libsdr_b200 import _lib
    from libsdr_b200.nodes import IQBaseBand, RxChain, DEMOD_FM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if args.gpus > 1 and world == 1:
        sys.stderr.write("bench.py: --gpus %d without torch.distributed.run (WORLD_SIZE unset): measuring ONE GPU, n_gpus=1 in the line\n" % args.gpus)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's version/info banner goes to stdout too, so it is
        # silenced unless asked for (SDRG_NCCL_DEBUG=INFO shows the NVLS/ring choice; not a timed run then)
        os.environ["NCCL_DEBUG"] = os.environ.get("SDRG_NCCL_DEBUG", "WARN")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    _lib.call("sdrg_set_device", local)
    dev = torch.device("cuda", local)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    c = workload()
    if args.buffers:
        c["n_buffers"] = args.buffers
    bs, nb = c["buffer_size"], c["n_buffers"]
    n_step = bs * nb
    x_host = torch.from_numpy(make_input(c)).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    bb = IQBaseBand("f32", c["Fc"], c["Ff"], c["width"], c["order"], c["sub_sample"], c["oFs"])
    bb.config(sample_rate=c["Fs"], buffer_size=bs)
    chain = RxChain(bb, DEMOD_FM)
    ss = bb.info().sub_sample
    n_out_cap = n_step // ss + 2
    bb_out = torch.empty((n_out_cap, 2), dtype=torch.float32, device=dev)
    # The demodulated audio of every step is gathered from all ranks with NCCL (the only collective on
    # this path).  It is latency-bound (2.6 MB per rank per step), so G steps are batched per
    # collective and the collective of one batch overlaps the kernels of the next (two rings).
    G = int(os.environ.get("SDRG_BENCH_G", "8"))
    rings = [torch.zeros((G, n_out_cap), dtype=torch.float32, device=dev) for _ in range(2)]
    gathered2 = [torch.empty((world, G, n_out_cap), dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    pending = [None, None]
    counter = [0]

    # Per-kernel CUDA events (sdrg_profile) are recorded on every 8th step of the timed region, the first one
    # included: the event records between the kernels cost ~3 % of the step when taken on every step.
    prof_every = int(os.environ.get("SDRG_BENCH_PROFILE_EVERY", "8"))
    prof_live = [False]
    prof_k0 = [0]
    prof_off = [0]                # offset inside each group of prof_every steps (away from the step that overlaps a gather launch)

    def step_dev():
        k = counter[0]
        counter[0] += 1
        if prof_live[0]:
            _lib.profile_enable((k - prof_k0[0]) % prof_every == prof_off[0])
        ring, slot = (k // G) & 1, k % G
        if slot == 0 and pending[ring] is not None:
            pending[ring].wait()                    # the gather that last read this ring has finished
            pending[ring] = None
        chain.process(x_dev, bs, bb_out=bb_out, audio_out=rings[ring][slot])
        if world > 1 and slot == G - 1:
            pending[ring] = dist.all_gather_into_tensor(gathered2[ring], rings[ring], async_op=True)

    def drain():
        if world > 1 and counter[0] % G:            # a partially filled ring at the end of the region
            ring = (counter[0] // G) & 1
            if pending[ring] is not None:
                pending[ring].wait()
            pending[ring] = dist.all_gather_into_tensor(gathered2[ring], rings[ring], async_op=True)
            counter[0] += G - counter[0] % G
        for i, w in enumerate(pending):
            if w is not None:
                w.wait()
                pending[i] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step_dev()
    drain()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.profile_read(_lib.KERNEL_IQBB_ACCUM); _lib.profile_read(_lib.KERNEL_IQBB_FINALIZE)
    prof_live[0] = True
    prof_k0[0] = counter[0]
    prof_off[0] = min(3, prof_every - 1) if K >= 4 else 0
    l0 = _lib.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step_dev()
    drain()
    ev1.record()
    barrier()
    launches = _lib.kernel_launch_count() - l0
    prof_live[0] = False
    _lib.profile_enable(False)
    ms = ev0.elapsed_time(ev1)
    acc_ms, acc_n = _lib.profile_read(_lib.KERNEL_IQBB_ACCUM)
    fin_ms, fin_n = _lib.profile_read(_lib.KERNEL_IQBB_FINALIZE)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n_step * K / (ms * 1e-3) / 1e6

    # ---- end to end: host pointers through the C ABI, pinned input, audio read back
    x_np = x_host.numpy()
    n_out = bb.outputs_for(n_step)
    audio_host = torch.zeros(n_out + 1, dtype=torch.float32).pin_memory().numpy()
    K2 = max(3, min(K, 5))

    def step_e2e():
        got = C.c_size_t(0)
        _lib.call("sdrg_rxchain_process", chain._h, C.c_void_p(x_np.ctypes.data), bs, nb, None,
                  C.c_void_p(audio_host.ctypes.data), n_out + 1, C.byref(got), None)
        return got.value

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    got = 0
    for _ in range(K2):
        got = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_step * K2 / e2e_s / 1e6
    clocks = sampler.stop() if rank == 0 else None
    secondary = secondary_workloads(world, rank, dev, dist, args)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        alg_bytes = n_step * 8 + (n_step // ss) * 4         # cf32 in + float FM out (8 + 4/ss B/sample)
        k_ms = acc_ms / max(acc_n, 1)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("iqbb_accum_f32_c2_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_block(c, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_step * 8),
                        "d2h_bytes_per_step": int(got * 4), "steps": K2},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "iqbb accumulate (FIR->NCO->window sums), float",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "kernel_timing": "CUDA events on the launching stream around the kernel, every %d-th step of the timed region (offset %d)" % (prof_every, prof_off[0]),
                             "peak_note": "the measured peak is a copy (read + write) figure; this kernel is a pure read stream, "
                                          "which HBM3e serves slightly faster, so frac can exceed 1",
                             "kernel_ms": k_ms, "kernel_launches": int(acc_n),
                             "kernel_share_of_step": (k_ms / (ms / K)) if ms > 0 and k_ms else None,
                             "finalize_ms": fin_ms / max(fin_n, 1)},
                "clocks": clocks, "secondary": secondary}
        if world == 1 and not args.no_cpu_baseline:
            nbuf = 24
            v, dt = cpu_port_run(c, x_np, 1, nbuf)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d buffers of %d samples, oracle port of the float chain (the reference "
                                              "has no float instantiation), %.1f s" % (nbuf, bs, dt)}
            ref = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
            if os.path.exists(ref):     # informational: the real reference on the int16 stand-in of the same shape
                try:
                    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
                        synth.iq_int(4 << 20, c["Fs"], [(8192, 103e3, 0.0), (4096, 99e3, 0.5)], 64, 1, np.int16).tofile(f)
                    out = subprocess.run([ref, "time", "s16", f.name, str(bs), repr(c["Fs"]), repr(c["Fc"]), repr(c["Ff"]),
                                          repr(c["width"]), str(c["order"]), "1", repr(c["oFs"]), "1", "2.0"],
                                         capture_output=True, text=True, timeout=120).stdout
                    os.unlink(f.name)
                    line["cpu_baseline"]["reference_int16_standin_msamples_per_s"] = json.loads(out)["msamples_per_s"]
                except Exception as e:  # pragma: no cover
                    line["cpu_baseline"]["reference_int16_standin_error"] = str(e)[:100]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
