#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

    python tests/golden/gen_golden.py

needs /root/reference (this container only): it builds oracle/_ref/ref_harness from the reference's
own sources (oracle/Makefile) and stores, per case, the input bytes, the node parameters the
reference derived (kernel, LUT, lut_inc, sub_sample) and every output it produced
(IQBaseBand, FMDemod in place, AMDemod, USBDemod; FilterSink+FilterSource) as compressed .npz.
The reference publishes no golden vectors of its own for this path.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libsdr_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

T3 = lambda a: [(a, 103e3, 0.0), (a / 2, 99e3, 0.5), (a / 3, 300e3, 1.0)]  # noqa: E731

# name, scalar, Fs, Fc, Ff, width, order, sub_sample, oFs, setcf, buffer_size, N, tone amplitude, noise
BB_CASES = [
    ("bb_s16_c1",        "s16", 2.4e6,  100e3,  100e3, 12.5e3, 15, 1, 48000.0, 0, 4096, 10007, 8192, 64),
    ("bb_s16_neg_full",  "s16", 2.4e6, -100e3, -100e3, 12.5e3, 21, 1, 48000.0, 0, 8192, 30000, 32767, 0),
    ("bb_s16_wrap",      "s16", 100e6,  100e3,  100e3, 25e3,   15, 1, 48000.0, 0, 16384, 50000, 8192, 64),
    ("bb_s16_sdrfm",     "s16", 1e6,    100e3,  100e3, 12.5e3, 21, 1, 8000.0,  1, 16384, 40000, 8192, 64),
    ("bb_s16_ss1",       "s16", 2.4e6,  100e3,  100e3, 12.5e3, 15, 1, 2.4e6,   0, 1024, 4099, 8192, 64),
    ("bb_s16_inc0",      "s16", 2.4e6,  0.0,    0.0,   12.5e3, 16, 1, 48000.0, 0, 8192, 30000, 8192, 64),
    ("bb_s16_ss7",       "s16", 2.4e6,  333e3,  250e3, 200e3,  33, 7, 0.0,     0, 1000, 9001, 12000, 64),
    ("bb_s16_fracfc",    "s16", 2.4e6,  100e3 + 0.75, 100e3, 12.5e3, 15, 1, 48000.0, 0, 4096, 9000, 8192, 64),
    ("bb_s8_pos",        "s8",  2.4e6,  100e3,  100e3, 12.5e3, 15, 1, 48000.0, 0, 4096, 20000, 60, 4),
    ("bb_s8_neg",        "s8",  2.4e6, -100e3, -100e3, 12.5e3, 21, 1, 48000.0, 0, 4096, 20000, 60, 4),
    ("bb_s8_inc0",       "s8",  2.4e6,  0.0,    0.0,   50e3,   16, 4, 0.0,     0, 2048, 10000, 60, 4),
    ("bb_s8_wrap",       "s8",  1e6,    100e3,  100e3, 400e3,  9,  2, 0.0,     0, 2048, 10000, 127, 0),
]

# name, block, Fs, fmin, fmax, number of blocks
OLA_CASES = [
    ("ola_n64",  64,  20e6, 100e3, 300e3, 12),
    ("ola_n256", 256, 20e6, -2e6, 1e6, 6),
    ("ola_n1024_swap", 1024, 20e6, 3e6, 1e6, 3),   # fmax < fmin handled by FilterNode::addFilter
]


def run(args):
    subprocess.run([HARNESS] + [str(a) for a in args], check=True)


def gen_cast_table(td):
    """The whole AutoCast table (src/autocast.hh:30-69) run by the reference on one random byte string: for every
    (source type, AutoCast<Out>) pair either the produced bytes or the fact that the reference refuses the pair."""
    g = np.random.Generator(np.random.MT19937(0xCA57))
    x = g.integers(0, 256, size=4096).astype(np.uint8)
    x[:16] = [0, 1, 126, 127, 128, 129, 254, 255, 0, 255, 255, 0, 127, 127, 128, 128]     # the edges of every width
    inp = os.path.join(td, "cast_in.bin"); x.tofile(inp)
    out = {"x": x, "buffer_bytes": 1024}
    refused = []
    for out_t in (2, 8, 4, 10):
        for in_t in range(1, 13):
            pre = os.path.join(td, "cast_%d_%d" % (in_t, out_t))
            r = subprocess.run([HARNESS, "castx", str(in_t), str(out_t), inp, "1024", pre])
            if r.returncode == 0:
                out["y_%d_%d" % (in_t, out_t)] = np.fromfile(pre + ".out", dtype=np.uint8)
            elif r.returncode == 4:
                refused.append((in_t, out_t))
            else:
                raise RuntimeError("harness failed on cast %d -> %d" % (in_t, out_t))
    out["refused"] = np.array(refused, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "cast_table.npz"), **out)
    print("cast_table: %d casts, %d refused pairs" % (len([k for k in out if k.startswith("y_")]), len(refused)))


def gen_rbb8(td):
    """Real-input BaseBand<int8_t> (src/baseband.hh:304-529 with Scalar = int8_t: 16-bit arithmetic throughout)."""
    for (name, Fs, Fc, Ff, width, order, ss, bs, N, amp) in [
            ("rbb8_pos", 48000.0, 10e3, 10e3, 3e3, 21, 6, 1000, 9000, 60),
            ("rbb8_neg_wrap", 2.4e6, -300.5e3, -280e3, 50e3, 32, 50, 4096, 30000, 127),     # full scale: every 16-bit wrap is exercised
            ("rbb8_inc0_ss1", 8000.0, 0.0, 1e3, 500.0, 9, 1, 512, 3000, 100),
            ("rbb8_ss200", 1e6, 123e3, 120e3, 20e3, 15, 200, 2048, 20000, 90)]:                 # ss^2 = 40000 wraps int16 in std::norm
        t = np.arange(N) / Fs
        x = (amp * 0.6 * np.cos(2 * np.pi * abs(Fc if Fc else 1e3) * 1.01 * t) + amp * 0.3 * np.cos(2 * np.pi * 0.37 * Fs / 2 * t + 1))
        x = np.clip(x + np.random.Generator(np.random.MT19937(0x5D12000D)).integers(-amp // 8, amp // 8 + 1, size=N), -128, 127).astype(np.int8)
        inp = os.path.join(td, name + ".in"); x.tofile(inp)
        pre = os.path.join(td, name)
        run(["rbb8", inp, bs, repr(Fs), repr(Fc), repr(Ff), repr(width), order, ss, pre])
        raw = np.fromfile(pre + ".params", dtype=np.uint8)
        hdr = raw[:32].view(np.int64); L = int(hdr[0])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), Fs=Fs, Fc=Fc, Ff=Ff, width=width, order=order, sub_sample=ss,
                            buffer_size=bs, x=x, ref_lut_inc=int(hdr[2]), ref_neg=int(hdr[3]),
                            ref_kernel=raw[32:32 + 8 * L].view(np.int32).reshape(L, 2),
                            bb=np.fromfile(pre + ".bb", dtype=np.int8).reshape(-1, 2),
                            counts=np.fromfile(pre + ".counts", dtype=np.uint32))
        print("wrote", name, "inc=%d outputs=%d" % (hdr[2], np.fromfile(pre + ".counts", dtype=np.uint32).sum()))


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if "--cast-only" in sys.argv or "--rbb8-only" in sys.argv:
        with tempfile.TemporaryDirectory() as td:
            gen_cast_table(td) if "--cast-only" in sys.argv else gen_rbb8(td)
        return
    with tempfile.TemporaryDirectory() as td:
        gen_cast_table(td)
        gen_rbb8(td)
        for (name, sc, Fs, Fc, Ff, width, order, ss, oFs, setcf, bs, N, amp, noise) in BB_CASES:
            dt = np.int16 if sc == "s16" else np.int8
            x = synth.iq_int(N, Fs, T3(amp), noise, 0x5D120000 + len(name), dt)
            inp = os.path.join(td, name + ".in"); x.tofile(inp)
            pre = os.path.join(td, name)
            run(["bb", sc, inp, bs, repr(Fs), repr(Fc), repr(Ff), repr(width), order, ss, repr(oFs), setcf, pre])
            raw = np.fromfile(pre + ".params", dtype=np.uint8)
            hdr = raw[:32].view(np.int64)
            L = int(hdr[0])
            kern = raw[32:32 + 8 * L].view(np.int32).reshape(L, 2)
            lut = raw[32 + 8 * L:32 + 8 * L + 8 * 128].view(np.int32).reshape(128, 2)
            np.savez_compressed(
                os.path.join(HERE, name + ".npz"),
                scalar=sc, Fs=Fs, Fc=Fc, Ff=Ff, width=width, order=order, sub_sample_arg=ss, oFs=oFs, setcf=setcf,
                buffer_size=bs, x=x,
                ref_order=L, ref_sub_sample=int(hdr[1]), ref_lut_inc=int(hdr[2]), ref_neg=int(hdr[3]),
                ref_kernel=kern, ref_lut=lut,
                bb=np.fromfile(pre + ".bb", dtype=dt).reshape(-1, 2),
                counts=np.fromfile(pre + ".counts", dtype=np.uint32),
                fm=np.fromfile(pre + ".fm", dtype=np.int16),
                am=np.fromfile(pre + ".am", dtype=dt),
                usb=np.fromfile(pre + ".usb", dtype=dt))
            print("wrote", name, "ss=%d inc=%d outputs=%d" % (hdr[1], hdr[2], np.fromfile(pre + ".counts", dtype=np.uint32).sum()))
        # real-input BaseBand<int16_t>: name, Fs, Fc, Ff, width, order, ss, buffer_size, N
        for (name, Fs, Fc, Ff, width, order, ss, bs, N) in [
                ("rbb_pos", 48000.0, 10e3, 10e3, 3e3, 21, 6, 1000, 9000),
                ("rbb_neg_even", 2.4e6, -300.5e3, -280e3, 50e3, 32, 50, 4096, 30000),
                ("rbb_inc0_ss1", 8000.0, 0.0, 1e3, 500.0, 9, 1, 512, 3000)]:
            t = np.arange(N) / Fs
            x = (9000 * np.cos(2 * np.pi * abs(Fc if Fc else 1e3) * 1.01 * t) + 4000 * np.cos(2 * np.pi * 0.37 * Fs / 2 * t + 1)).astype(np.int16)
            x += np.random.Generator(np.random.MT19937(0x5D12000B)).integers(-200, 201, size=N).astype(np.int16)
            inp = os.path.join(td, name + ".in"); x.tofile(inp)
            pre = os.path.join(td, name)
            run(["rbb", inp, bs, repr(Fs), repr(Fc), repr(Ff), repr(width), order, ss, pre])
            raw = np.fromfile(pre + ".params", dtype=np.uint8)
            hdr = raw[:32].view(np.int64); L = int(hdr[0])
            np.savez_compressed(os.path.join(HERE, name + ".npz"), Fs=Fs, Fc=Fc, Ff=Ff, width=width, order=order, sub_sample=ss,
                                buffer_size=bs, x=x, ref_lut_inc=int(hdr[2]), ref_neg=int(hdr[3]),
                                ref_kernel=raw[32:32 + 8 * L].view(np.int32).reshape(L, 2),
                                bb=np.fromfile(pre + ".bb", dtype=np.int16).reshape(-1, 2),
                                counts=np.fromfile(pre + ".counts", dtype=np.uint32))
            print("wrote", name, "inc=%d" % hdr[2])
        # WavSink<T> -> file -> WavSource: the file bytes the reference writes and the buffers it reads back
        gw = np.random.Generator(np.random.MT19937(0x5D12000C))
        for typ, dt, ncomp, Fs, bs, frames in (("u8", np.uint8, 1, 8000.0, 512, 1300), ("s16", np.int16, 1, 48000.0, 1024, 2500),
                                               ("cu8", np.uint8, 2, 1e6, 1024, 3000), ("cs16", np.int16, 2, 2.4e6, 2048, 4096)):
            x = gw.integers(np.iinfo(dt).min, np.iinfo(dt).max + 1, size=frames * ncomp).astype(dt)
            inp = os.path.join(td, "wav_" + typ + ".in"); x.tofile(inp)
            pre = os.path.join(td, "wav_" + typ); wav = pre + ".wav"
            run(["wav", typ, inp, bs, repr(Fs), wav, pre])
            np.savez_compressed(os.path.join(HERE, "wav_" + typ + ".npz"), x=x, Fs=Fs, buffer_size=bs,
                                wav=np.fromfile(wav, dtype=np.uint8), data=np.fromfile(pre + ".data", dtype=np.uint8),
                                byte_counts=np.fromfile(pre + ".counts", dtype=np.uint32), cfg=np.fromfile(pre + ".cfg"))
            print("wrote wav_" + typ)
        # AutoCast<complex<int16>> from cu8 / cs8, and FMDeemph<int16> (the nodes either side of the path)
        g = np.random.Generator(np.random.MT19937(0x5D12000A))
        for fmt, dt in (("cu8", np.uint8), ("cs8", np.int8)):
            x = g.integers(np.iinfo(dt).min, np.iinfo(dt).max + 1, size=(5000, 2)).astype(dt)
            x[:4] = [[0, 255], [127, 128], [1, 254], [126, 129]] if dt == np.uint8 else [[-128, 127], [0, -1], [1, -127], [126, -126]]
            inp = os.path.join(td, "cast_" + fmt + ".in"); x.tofile(inp)
            pre = os.path.join(td, "cast_" + fmt)
            run(["cast", fmt, inp, 1024, pre])
            np.savez_compressed(os.path.join(HERE, "cast_" + fmt + ".npz"), x=x,
                                out=np.fromfile(pre + ".cs16", dtype=np.int16).reshape(-1, 2))
            print("wrote cast_" + fmt)
        for name, Fs in (("deemph_48k", 48000.0), ("deemph_8k", 8000.0), ("deemph_200k", 200e3)):
            x = (8000 * np.sin(2 * np.pi * 1e3 * np.arange(6000) / Fs)).astype(np.int16) + g.integers(-3000, 3001, size=6000).astype(np.int16)
            x[100:110] = [32767, -32768, 32767, -32768, 0, 1, -1, 2, -2, 32767]
            inp = os.path.join(td, name + ".in"); x.tofile(inp)
            pre = os.path.join(td, name)
            run(["deemph", inp, 1000, repr(Fs), pre])
            np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, Fs=Fs, buffer_size=1000,
                                out=np.fromfile(pre + ".out", dtype=np.int16))
            print("wrote", name)
        for (name, block, Fs, fmin, fmax, nblk) in OLA_CASES:
            x = synth.iq_f32(block * nblk, Fs, [(0.5, 200e3, 0.0), (0.25, 1.5e6, 0.5), (0.1667, -3e6, 1.0)], 0.01, 0x5D120003)
            inp = os.path.join(td, name + ".in"); x.tofile(inp)
            pre = os.path.join(td, name)
            lo, hi = (fmin, fmax) if fmin <= fmax else (fmax, fmin)   # FilterNode::addFilter swaps
            run(["ola", inp, block, repr(Fs), repr(lo), repr(hi), pre])
            np.savez_compressed(
                os.path.join(HERE, name + ".npz"),
                block=block, Fs=Fs, fmin=fmin, fmax=fmax, x=x,
                kern=np.fromfile(pre + ".kern", dtype=np.complex64),
                taps=np.fromfile(pre + ".taps", dtype=np.complex64),
                out=np.fromfile(pre + ".out", dtype=np.complex64))
            print("wrote", name)


if __name__ == "__main__":
    main()
