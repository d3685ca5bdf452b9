"""numpy model of the window-direct float kernel's algebra (iqbb_fold_direct.cu) against the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc


def model(x, k, lut, inc, neg, ss, consumed0=0):
    """whole-stream: out[w] for windows whose halo is inside x; x complex128, first sample = stream start"""
    L = len(k); L1 = L - 1; ln = ss + L1
    n = len(x)
    # class table
    rows = []; cls = np.zeros(256, dtype=int); prev = None
    for r in range(256):
        sig = tuple(((r + d * inc) >> 8) for d in range(ss)) if inc else tuple([0] * ss)
        if sig != prev:
            B = np.array([(lut[(127 - (s % 128)) % 128] if neg else lut[s % 128]) if inc else 1.0 for s in sig])
            V = np.zeros(ln, dtype=complex)
            for j in range(ln):
                for d in range(max(0, j - L1), min(j, ss - 1) + 1):
                    V[j] += B[d] * k[j - d]
            rows.append(V); prev = sig
        cls[r] = len(rows) - 1
    A = np.conj(lut) if neg else lut
    out = {}
    s = 1
    while True:
        nb = s * ss + 1                 # first=1, r0=0
        if nb + ss > n: break
        if nb - L1 >= 0:
            ph = (nb * inc) & 0x7fff
            a = A[ph >> 8] if inc else 1.0
            out[s] = a * np.dot(rows[cls[ph & 255]], x[nb - L1: nb + ss]) / ss
        s += 1
    return out, len(rows)


rng = np.random.default_rng(1)
for (Fc, order, ss) in [(100e3, 15, 16), (-100e3, 15, 50), (100e3, 32, 64), (0.0, 9, 8), (333e3, 64, 63), (100e3, 15, 14), (-7e5, 5, 2)]:
    o = orc.IQBaseBand(orc.F32, Fc, 100e3, 12.5e3, order, ss, 0.0)
    o.config(20e6, 1 << 16)
    n = 4096
    x = rng.standard_normal((n, 2)).astype(np.float32)
    ref = o.process(x)
    ref = ref[:, 0].astype(float) + 1j * ref[:, 1]
    k = o.kernel_f64()
    lut = np.exp(-2j * np.pi * np.arange(128) / 128)
    got, nrows = model(x[:, 0].astype(float) + 1j * x[:, 1], k, lut, o.lut_inc & 0x7fff, o.neg, ss)
    err = max(abs(got[s] - ref[s]) for s in got) / np.sqrt(np.mean(abs(ref) ** 2))
    print("Fc=%g L=%d ss=%d inc=%d rows=%d windows=%d max rel err %.2e" % (Fc, order, ss, o.lut_inc, nrows, len(got), err))
