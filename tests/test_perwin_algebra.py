"""CPU check of the algebra behind the per-window float kernel (libsdr_b200/csrc/iqbb_fold_perwin.cu):
out[w] = A(a_b) * sum_j V(r_b, j) x[n_b - (L-1) + j] with V(r, .) depending on r only through <= ss carry patterns.
A numpy restatement of the table construction (api.cu upload_fold_tables) and of the window sums is compared with the
oracle's sample-serial float chain; the CUDA kernel itself is covered by tests/test_gpu_float_perwin.py."""
import numpy as np
import pytest

from oracle import oracle as orc


def v_table(k, lut, inc, neg, ss):
    """rows: one V(r, .) per distinct carry pattern; cls[r] -> row (monotone in r)"""
    L1 = len(k) - 1
    rows, cls, prev = [], np.zeros(256, dtype=int), None
    for r in range(256):
        sig = tuple((r + d * inc) >> 8 for d in range(ss)) if inc else (0,) * ss
        if sig != prev:
            B = np.array([(lut[(127 - s) % 128] if neg else lut[s % 128]) if inc else 1.0 for s in sig], dtype=complex)
            V = np.zeros(ss + L1, dtype=complex)
            for j in range(ss + L1):
                for d in range(max(0, j - L1), min(j, ss - 1) + 1):
                    V[j] += B[d] * k[j - d]
            rows.append(V)
            prev = sig
        cls[r] = len(rows) - 1
    return rows, cls


@pytest.mark.parametrize("Fc,order,ss", [(100e3, 15, 16), (-100e3, 15, 50), (100e3, 32, 64), (0.0, 9, 8), (333e3, 64, 63),
                                         (100e3, 15, 14), (-7e5, 5, 2), (1.25e6, 20, 128)])
def test_window_sums_with_deduplicated_rows_match_the_oracle(Fc, order, ss):
    o = orc.IQBaseBand(orc.F32, Fc, 100e3, 12.5e3, order, ss, 0.0)
    o.config(20e6, 1 << 16)
    n = 3000
    x = np.random.default_rng(1).standard_normal((n, 2)).astype(np.float32)
    ref = o.process(x)
    ref = ref[:, 0].astype(float) + 1j * ref[:, 1]
    xc = x[:, 0].astype(float) + 1j * x[:, 1]
    k, inc, neg = o.kernel_f64(), o.lut_inc & 0x7fff, o.neg
    lut = np.exp(-2j * np.pi * np.arange(128) / 128)
    rows, cls = v_table(k, lut, inc, neg, ss)
    assert len(rows) <= max(ss, 1)                       # at most ss distinct rows, not 256
    assert np.all(np.diff(cls) >= 0)                     # classes are intervals of r
    A = np.conj(lut) if neg else lut
    L1 = order - 1
    checked = 0
    s = 1
    while (s + 1) * ss + 1 <= n:                         # window s = samples [s ss + 1, (s+1) ss + 1) of a stream that starts here
        nb = s * ss + 1
        if nb - L1 >= 0:
            ph = (nb * inc) & 0x7fff
            got = (A[ph >> 8] if inc else 1.0) * np.dot(rows[cls[ph & 255]], xc[nb - L1: nb + ss]) / ss
            assert abs(got - ref[s]) <= 3e-7 * max(1.0, np.sqrt(np.mean(np.abs(ref) ** 2))) + 1e-6 * abs(ref[s])
            checked += 1
        s += 1
    assert checked > 10
