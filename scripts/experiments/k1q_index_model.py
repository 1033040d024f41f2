"""CPU model of the k1q kernel's index arithmetic (blind_rotate_k1q.cu): the three-pass (RA x 16 x 4) negacyclic
transform in position order, its inverse, the shared-memory swizzle and the resident-key addressing.  Run on the
CPU box before spending GPU time: every assertion here is a property the CUDA code relies on."""
import numpy as np


def brev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def swz(idx):
    return idx ^ ((idx >> 3) & 7) ^ (((idx >> 6) & 1) << 2)


def c_of_thread(tid):
    """pass C / C': thread -> position group c (positions 4c..4c+3).  Lanes 0..15 of warp w take the even groups of
    key columns c' = 16w + lane, lanes 16..31 the odd ones, so that the 8 lanes of a quarter warp read 8 consecutive
    16-byte key words (ONE 128-byte line per quarter-warp request; with c = tid they straddle two lines and the key
    loads cost twice the L1 wavefronts: ncu r2a)."""
    w, lane = tid >> 5, tid & 31
    return 2 * (16 * w + (lane & 15)) + (lane >> 4)


def dif(x):
    """reg_dif: out[pos] = X[brev(pos)], X_k = sum_m x_m W_R^(+mk)"""
    R = len(x)
    lg = R.bit_length() - 1
    X = np.array([sum(x[m] * np.exp(2j * np.pi * m * k / R) for m in range(R)) for k in range(R)])
    return np.array([X[brev(p, lg)] for p in range(R)])


def dit_inv(xp):
    """reg_dit_inv: in[pos] = X[brev(pos)], out[m] = sum_k X_k W_R^(-mk)"""
    R = len(xp)
    lg = R.bit_length() - 1
    X = np.zeros(R, complex)
    for p in range(R):
        X[brev(p, lg)] = xp[p]
    return np.array([sum(X[k] * np.exp(-2j * np.pi * m * k / R) for k in range(R)) for m in range(R)])


def forward(a, N):
    M = N // 2
    RA = M // 64
    lgA = RA.bit_length() - 1
    w = np.exp(1j * np.pi / N)
    row = np.zeros(M, complex)           # physical (swizzled) storage
    # pass A: thread q < 64
    for q in range(64):
        x = np.array([(a[q + 64 * m] + 1j * a[q + 64 * m + M]) * w ** (64 * m) for m in range(RA)])
        x = dif(x)
        for pos1 in range(RA):
            k1 = brev(pos1, lgA)
            tw = np.exp(1j * np.pi * q * (4 * k1 + 1) / N)
            row[swz(pos1 * 64 + q)] = x[pos1] * tw
    # pass B: task (pos1, r)
    for pos1 in range(RA):
        for r in range(4):
            y = np.array([row[swz(pos1 * 64 + r + 4 * m2)] for m2 in range(16)])
            y = dif(y)
            for pos2 in range(16):
                row[swz(pos1 * 64 + pos2 * 4 + r)] = y[pos2]
    # pass C: thread c < M/4
    out = np.zeros(M, complex)
    for c in range(M // 4):
        pos2 = c % 16
        k2 = brev(pos2, 4)
        v = np.array([row[swz(4 * c + r)] * np.exp(2j * np.pi * r * k2 / 64) for r in range(4)])
        v = dif(v)
        for pos3 in range(4):
            out[4 * c + pos3] = v[pos3]
    return out


def inverse(Y, N):
    M = N // 2
    RA = M // 64
    lgA = RA.bit_length() - 1
    w = np.exp(1j * np.pi / N)
    row = np.zeros(M, complex)
    for c in range(M // 4):
        pos2 = c % 16
        k2 = brev(pos2, 4)
        v = dit_inv(np.array([Y[4 * c + p] for p in range(4)]))
        for r in range(4):
            row[swz(4 * c + r)] = v[r] * np.exp(-2j * np.pi * r * k2 / 64)
    for pos1 in range(RA):
        for r in range(4):
            y = dit_inv(np.array([row[swz(pos1 * 64 + pos2 * 4 + r)] for pos2 in range(16)]))
            for m2 in range(16):
                row[swz(pos1 * 64 + r + 4 * m2)] = y[m2]
    a = np.zeros(N)
    for q in range(64):
        x = np.zeros(RA, complex)
        for pos1 in range(RA):
            k1 = brev(pos1, lgA)
            x[pos1] = row[swz(pos1 * 64 + q)] * np.exp(-1j * np.pi * q * (4 * k1 + 1) / N)
        x = dit_inv(x)
        for m in range(RA):
            z = x[m] * w ** (-64 * m) / M
            a[q + 64 * m] = z.real
            a[q + 64 * m + M] = z.imag
    return a


def check_conflicts(M):
    """every 128-bit shared-memory access: the 8 threads of a quarter warp must hit 8 distinct 16-byte bank groups"""
    RA = M // 64

    def ok(idxs):
        return len({swz(i) & 7 for i in idxs}) == 8

    for qw in range(8):                                   # pass A / A': 8 consecutive q, any pos1
        for pos1 in range(RA):
            assert ok([pos1 * 64 + 8 * qw + i for i in range(8)])
    for w0 in range(0, RA * 4, 8):                        # pass B / B': lanes (row, pos1 & 1, r), pos1 = 2*warp + bit
        tasks = [(t // 4, t % 4) for t in range(w0, w0 + 8)]
        for m2 in range(16):
            assert ok([p1 * 64 + r + 4 * m2 for p1, r in tasks])
            assert ok([p1 * 64 + m2 * 4 + r for p1, r in tasks])
    for t0 in range(0, M // 4, 8):                        # pass C / C'
        cs = [c_of_thread(t0 + i) for i in range(8)]
        for r in range(4):
            assert ok([4 * c + r for c in cs])
        # key loads: position 4c + i = 8c' + m' is stored at m' * (M/8) + c': one 128-byte line per quarter warp
        for i in range(4):
            words = [(4 * (c & 1) + i) * (M // 8) + (c >> 1) for c in cs]
            assert max(words) - min(words) == 7 and min(words) % 8 == 0, words
    # pass B / B' run per warp on the blocks pos1 in {2w, 2w+1} that the same warp's pass C / C' threads own
    for t in range(M // 4):
        assert (c_of_thread(t) >> 4) >> 1 == t >> 5


for N in (512, 1024, 2048):
    M = N // 2
    lg = M.bit_length() - 1
    rng = np.random.default_rng(N)
    a = rng.integers(-32, 32, N).astype(float)
    X = forward(a, N)
    w = np.exp(1j * np.pi / N)
    z = (a[:M] + 1j * a[M:]) * w ** np.arange(M)
    for s in range(M):                                    # position s holds the value at root exponent 1 + 4*brev(s)
        k = brev(s, lg)
        want = sum(z * np.exp(2j * np.pi * np.arange(M) * k / M))
        assert abs(X[s] - want) < 1e-7 * max(1, abs(want)), (N, s)
        # = the polynomial evaluated at w^(1+4k)
    back = inverse(X, N)
    assert np.abs(back - a).max() < 1e-9, N
    check_conflicts(M)
    print(f"N={N}: forward positions, inverse round trip and bank-conflict freedom OK")


# ---- the lane-pair split of a radix-16 task (two-row phases of blind_rotate_k1q.cu) ---------------------------------------
def dif16_pair(x):
    out = np.zeros(16, complex)
    for h in (0, 1):
        sg = -1.0 if h else 1.0
        y = np.array([x[i] + sg * x[i + 8] for i in range(8)])
        if h:
            y = y * np.exp(2j * np.pi * np.arange(8) / 16)
        out[8 * h: 8 * h + 8] = dif(y)
    return out


def dit16_pair(xp):
    halves = []
    for h in (0, 1):
        y = dit_inv(xp[8 * h: 8 * h + 8])
        if h:
            y = y * np.exp(-2j * np.pi * np.arange(8) / 16)
        halves.append(y)
    out = np.zeros(16, complex)
    for h in (0, 1):
        sg = -1.0 if h else 1.0
        mine, recv = halves[h], halves[1 - h]
        out[8 * h: 8 * h + 8] = recv + sg * mine if h else mine + recv
    return out


rng = np.random.default_rng(5)
v = rng.normal(size=16) + 1j * rng.normal(size=16)
assert np.abs(dif16_pair(v) - dif(v)).max() < 1e-12
assert np.abs(dit16_pair(v) - dit_inv(v)).max() < 1e-12
print("lane-pair split of the radix-16 task OK")


# ---- the folded decimation-in-time networks of k1_common.cuh (reg_dft_fma / reg_dit_inv_fma): index algebra only ----------
def w64(idx):
    return np.exp(2j * np.pi * (idx % 64) / 64)


def reg_dft_fma(x, TW):
    R = len(x)
    lg = R.bit_length() - 1
    a = np.array([x[brev(i, lg)] for i in range(R)], complex)
    h = 1
    while h < R:
        for b in range(R // 2):
            j = b & (h - 1)
            i0 = ((b - j) << 1) + j
            i1 = i0 + h
            w = w64(TW * (R // (2 * h)) + j * (32 // h))
            a[i0], a[i1] = a[i0] + w * a[i1], a[i0] - w * a[i1]
        h <<= 1
    return np.array([a[brev(i, lg)] for i in range(R)])


for R, TW in ((4, 0), (8, 0), (16, 0), (8, 2), (16, 1), (8, 5)):
    lg = R.bit_length() - 1
    v = rng.normal(size=R) + 1j * rng.normal(size=R)
    got = reg_dft_fma(v, TW)
    t = w64(TW)
    for pos in range(R):
        k = brev(pos, lg)
        want = sum(v[m] * (t * np.exp(2j * np.pi * k / R)) ** m for m in range(R))
        assert abs(got[pos] - want) < 1e-12, (R, TW, pos)
    if TW == 0:
        assert np.abs(got - dif(v)).max() < 1e-12
print("folded DIT networks (twist absorbed) OK")
