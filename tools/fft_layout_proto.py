"""numpy emulation of the per-warp 512-point complex FFT dataflow used by csrc/features.cu
(32 lanes x 16 complex registers, three radix-8 passes, two shared-memory transposes).
Run: python tools/fft_layout_proto.py  -- asserts equality with numpy.fft and bank-conflict freedom."""
import numpy as np

N = 512
rng = np.random.default_rng(0)
z = rng.standard_normal(N) + 1j * rng.standard_normal(N)
W = lambda n, e: np.exp(-2j * np.pi * e / n)

def dft8(v):  # v[8] -> DFT_8
    k = np.arange(8)
    return np.array([(v * W(8, k * kk)).sum() for kk in range(8)])

def banks_ok(idx_by_lane):
    # 8-byte (float2) elements: a warp access is served as two half-warp wavefronts; conflict free iff the 16 lanes of
    # each half hit 16 distinct 8-byte bank pairs (idx mod 16)
    idx = np.array(idx_by_lane)
    return all(len(set((idx[h * 16:(h + 1) * 16] % 16).tolist())) == 16 for h in range(2))

S1 = 72   # exchange-1 layout: idx = 72*k2 + 8*n1 + n0
S0 = 66   # exchange-2 layout: idx = 66*n0 + 8*k1 + k2 (float2 elements)
buf = np.zeros(576, complex)

# ---- stage 1: lane l, half h: butterfly g = l + 32h = n0 + 8 n1 ; inputs z[g + 64 n2]
regs = np.zeros((32, 2, 8), complex)
for l in range(32):
    for h in range(2):
        g = l + 32 * h
        n0, n1 = g % 8, g // 8
        a = dft8(np.array([z[g + 64 * n2] for n2 in range(8)]))
        regs[l, h] = a * W(64, n1 * np.arange(8))      # twiddle W64^{n1 k2}
# exchange 1 write (slot k2, half h): check banks across lanes
for h in range(2):
    for k2 in range(8):
        idxs = []
        for l in range(32):
            g = l + 32 * h; n0, n1 = g % 8, g // 8
            i = S1 * k2 + 8 * n1 + n0
            buf[i] = regs[l, h, k2]; idxs.append(i)
        assert banks_ok(idxs)
# ---- stage 2: lane l, half h: (n0 = l%8, k2 = l//8 + 4h); reads n1 = 0..7
regs2 = np.zeros((32, 2, 8), complex)
for h in range(2):
    for n1 in range(8):
        idxs = [S1 * (l // 8 + 4 * h) + 8 * n1 + l % 8 for l in range(32)]
        assert banks_ok(idxs)
for l in range(32):
    for h in range(2):
        n0, k2 = l % 8, l // 8 + 4 * h
        v = np.array([buf[S1 * k2 + 8 * n1 + n0] for n1 in range(8)])
        b = dft8(v)
        k1 = np.arange(8)
        regs2[l, h] = b * W(512, n0 * (k2 + 8 * k1))   # twiddle W512^{n0 (k2 + 8 k1)}
buf2 = np.zeros(576, complex)
for h in range(2):
    for k1 in range(8):
        idxs = []
        for l in range(32):
            n0, k2 = l % 8, l // 8 + 4 * h
            i = S0 * n0 + 8 * k1 + k2
            buf2[i] = regs2[l, h, k1]; idxs.append(i)
        assert banks_ok(idxs)
# ---- stage 3: lane l, half h: j = l + 32h = k2 + 8 k1 ; reads n0 = 0..7 ; outputs X[j + 64 k0]
X = np.zeros(N, complex)
for h in range(2):
    for n0 in range(8):
        idxs = []
        for l in range(32):
            j = l + 32 * h; k2, k1 = j % 8, j // 8
            idxs.append(S0 * n0 + 8 * k1 + k2)
        assert banks_ok(idxs)
for l in range(32):
    for h in range(2):
        j = l + 32 * h; k2, k1 = j % 8, j // 8
        v = np.array([buf2[S0 * n0 + 8 * k1 + k2] for n0 in range(8)])
        c = dft8(v)
        for k0 in range(8):
            X[j + 64 * k0] = c[k0]
ref = np.fft.fft(z)
print("max err", np.abs(X - ref).max())
assert np.abs(X - ref).max() < 1e-10
# two real frames packed: z = x1 + i x2
x1, x2 = rng.standard_normal(N), rng.standard_normal(N)
Z = np.fft.fft(x1 + 1j * x2)
k = np.arange(257); Zc = np.conj(Z[(N - k) % N])
P1 = np.abs((Z[k] + Zc) / 2) ** 2; P2 = np.abs((Z[k] - Zc) / 2j) ** 2
assert np.allclose(P1, np.abs(np.fft.rfft(x1)) ** 2) and np.allclose(P2, np.abs(np.fft.rfft(x2)) ** 2)
print("ok; max buffer index", max(S1 * 7 + 63, S0 * 7 + 63))
