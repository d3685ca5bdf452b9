"""numpy model of the index / twiddle algebra of csrc/fft8k_kernels.cu (no GPU needed): the DIF 2x(16x16x16)
forward pass, the permuted spectrum multiply, the DIT inverse and the overlap-save combine, checked against
numpy.fft.  Thread loops are vectorised over t = 0..255."""
import numpy as np

N = 4096
t = np.arange(256)
w = lambda n, k: np.exp(-2j * np.pi * (np.asarray(k) % n) / n)
T1 = w(4096, np.outer(np.arange(16), np.arange(256)))      # [q][u]
T2 = w(256, np.outer(np.arange(16), np.arange(16)))        # [q][c]
T8 = w(8192, np.arange(256)); T32 = w(32, np.arange(16))
dft16 = lambda v, inv: (np.fft.ifft(v, axis=0) * 16 if inv else np.fft.fft(v, axis=0))   # v[16, threads]
cm = lambda a, tw, inv: a * (np.conj(tw) if inv else tw)


def dif_pass(e, o, inv, conv=False):
    """e, o: [16, 256] register files (index a, thread t = 16 b + c). Returns U_h at [qc, t3], t3 = qa + 16 qb."""
    H = [np.zeros(4096, complex), np.zeros(4096, complex)]
    for h, v in enumerate((e, o)):
        v = dft16(v, inv)
        for q in range(16):
            H[h][256 * q + t] = cm(v[q], T1[q, t], inv)
    out = []
    for h in range(2):
        qa, c = t >> 4, t & 15
        v = np.stack([H[h][256 * qa + 16 * b + c] for b in range(16)])
        v = dft16(v, inv)
        for q in range(16):
            H[h][256 * qa + 16 * q + c] = cm(v[q], T2[q, c], inv)
    for h in range(2):
        qa, qb = (t >> 4, t & 15) if conv else (t & 15, t >> 4)      # conv: warp-local mapping, bins qa + 16 qb + 256 qc
        v = np.stack([H[h][256 * qa + 16 * qb + cc] for cc in range(16)])
        out.append(dft16(v, inv))          # [qc, t3] = U_h[qa + 16 qb + 256 qc]
    return out


def fft8192(x, inv=False):
    p, c = x[:4096], x[4096:]
    a = np.arange(16)[:, None]
    j = 256 * a + t[None, :]
    e = p[j] + c[j]
    o = cm(p[j] - c[j], T8[t][None, :] * T32[a], inv)
    E, O = dif_pass(e, o, inv)
    X = np.zeros(8192, complex)
    for qc in range(16):
        X[2 * (t + 256 * qc)] = E[qc]
        X[2 * (t + 256 * qc) + 1] = O[qc]
    return X


def conv_block(prev, cur, K):
    a = np.arange(16)[:, None]
    j = 256 * a + t[None, :]
    e = prev[j] + cur[j]
    o = (prev[j] - cur[j]) * (T8[t][None, :] * T32[a])
    U = dif_pass(e, o, False, conv=True)
    Kp = np.zeros((2, 16, 256), complex)
    for h in range(2):
        for qc in range(16):
            Kp[h, qc] = K[2 * ((t >> 4) + 16 * (t & 15) + 256 * qc) + h] / 8192
    H = [np.zeros(4096, complex), np.zeros(4096, complex)]
    for h in range(2):
        qa, qb = t >> 4, t & 15
        v = dft16(U[h] * Kp[h], True)                       # index c'
        for cc in range(16):
            H[h][256 * qa + 16 * qb + cc] = cm(v[cc], T2[cc, qb], True)
    for h in range(2):
        qa, c = t >> 4, t & 15
        v = dft16(np.stack([H[h][256 * qa + 16 * qb + c] for qb in range(16)]), True)     # index b
        for b in range(16):
            H[h][256 * qa + 16 * b + c] = cm(v[b], T1[qa, 16 * b + c], True)
    vv = []
    for h in range(2):
        vv.append(dft16(np.stack([H[h][256 * qa + t] for qa in range(16)]), True))         # index a -> v_h[256 a + t]
    y = vv[0] - np.conj(T8[t][None, :] * T32[a]) * vv[1]
    out = np.zeros(4096, complex)
    out[j] = y
    return out


g = np.random.default_rng(1)
x = g.standard_normal(8192) + 1j * g.standard_normal(8192)
print("fwd", np.abs(fft8192(x) - np.fft.fft(x)).max())
print("inv", np.abs(fft8192(x, True) - np.fft.ifft(x) * 8192).max())
h = g.standard_normal(4096) + 1j * g.standard_normal(4096)
K = np.fft.fft(np.concatenate([h, np.zeros(4096)]))
prev, cur = x[:4096], x[4096:]
ref = np.convolve(np.concatenate([prev, cur]), h)[4096:8192]
print("conv", np.abs(conv_block(prev, cur, K) - ref).max() / np.abs(ref).max())
