"""np_spec.py — a SECOND, independent restatement of the gate-bootstrapping spec in plain numpy
integer arithmetic.  TEST INFRASTRUCTURE ONLY (imported by tests/ alone): it exists to pin the C oracle
(oracle/tfhe_oracle.c) and the committed golden vectors against something that shares no code with them.

Everything is exact: Torus32 values are Python/numpy integers reduced mod 2^32, the negacyclic products are O(N^2)
integer convolutions in int64 (digits are < 2^7, key words < 2^31, 6 rows x 1024 terms: every partial sum fits
comfortably in 63 bits).  Slow by design (about 15 s for one n = 630 bootstrap).  Each function cites the spec step;
the reference call sites those steps serve are src/Math.cpp:34-43 (boots* gates) and :261,277 (bootsMUX).
"""
import numpy as np

M32 = (1 << 32) - 1
EIGHTH = 1 << 29

# gate constants (SURVEY.md §8a, row "bootsAND/OR/XOR/XNOR"): t = (0, c0/8) + sa*ca + sb*cb
GATE = {"NAND": (1, -1, -1), "AND": (-1, 1, 1), "OR": (1, 1, 1), "XOR": (2, 2, 2), "XNOR": (-2, -2, -2),
        "NOR": (-1, -1, -1), "ANDNY": (-1, -1, 1), "ANDYN": (-1, 1, -1), "ORNY": (1, -1, 1), "ORYN": (1, 1, -1)}


def u32(x):
    return np.asarray(x).astype(np.int64) & M32


def prologue(op, ca, cb):
    """step 1"""
    c0, sa, sb = GATE[op]
    t = (sa * u32(ca) + sb * u32(cb)) & M32
    t[-1] = (t[-1] + c0 * EIGHTH) & M32
    return t


def mod_switch(t, N=1024):
    """step 2: round each word to the nearest multiple of 2^-log2(2N); result in [0, 2N)"""
    shift = 32 - (N.bit_length())          # 2N = 2^(log2 N + 1); N.bit_length() == log2 N + 1
    return ((u32(t) + (1 << (shift - 1))) >> shift) & (2 * N - 1)


def rotate(p, a, N=1024):
    """X^a * p mod (X^N + 1), 0 <= a < 2N"""
    p = u32(p)
    out = np.empty(N, dtype=np.int64)
    for j in range(N):  # explicit on purpose: no clever index arithmetic shared with the oracle
        src = j - a
        sign = 1
        while src < 0:
            src += N
            sign = -sign
        out[j] = (sign * p[src]) & M32
    return out


def negacyclic(d, k, N=1024):
    """d (small signed integers) times k (Torus32 words as signed int32) mod (X^N + 1), mod 2^32"""
    full = np.convolve(d.astype(np.int64), k.astype(np.int64))      # exact: |sum| < 2^48
    res = full[:N].copy()
    res[: N - 1] -= full[N:]
    return res & M32


def decompose(x, l, Bgbit):
    """step 4: l signed digits of every coefficient (offset trick), most significant first"""
    half = 1 << (Bgbit - 1)
    offset = sum(half << (32 - p * Bgbit) for p in range(1, l + 1)) & M32
    v = (u32(x) + offset) & M32
    return [((v >> (32 - p * Bgbit)) & ((1 << Bgbit) - 1)) - half for p in range(1, l + 1)]


def cmux(acc, bk_i, abar, l, Bgbit, N=1024):
    """step 4 for one mask element: ACC += BK_i (x) ((X^abar - 1) ACC); acc = (mask, body), bk_i [2l][2][N] int32"""
    if abar == 0:
        return [acc[0].copy(), acc[1].copy()]
    t = [(rotate(acc[c], abar, N) - u32(acc[c])) & M32 for c in range(2)]
    digits = decompose(t[0], l, Bgbit) + decompose(t[1], l, Bgbit)   # rows 0..l-1 from the mask, l..2l-1 from the body
    out = []
    for c in range(2):
        s = np.zeros(N, dtype=np.int64)
        for r in range(2 * l):
            s = (s + negacyclic(digits[r], bk_i[r][c], N)) & M32
        out.append((u32(acc[c]) + s) & M32)
    return out


def bootstrap_woks(t, bk, l, Bgbit, N=1024):
    """steps 2-5: t (n+1 words) -> extracted sample (N+1 words); bk [n][2l][2][N] int32"""
    n = len(t) - 1
    bar = mod_switch(t, N)
    mu = np.full(N, EIGHTH, dtype=np.int64)
    acc = [np.zeros(N, dtype=np.int64), rotate(mu, int((2 * N - bar[n]) % (2 * N)), N)]     # step 3
    for i in range(n):
        acc = cmux(acc, bk[i], int(bar[i]), l, Bgbit, N)
    u = np.empty(N + 1, dtype=np.int64)                                                       # step 5
    u[0] = acc[0][0]
    for j in range(1, N):
        u[j] = (-acc[0][N - j]) & M32
    u[N] = acc[1][0]
    return u


def keyswitch(u, ksk, t_digits=8, basebit=2):
    """step 6: u (N+1 words) -> n+1 words; ksk [N][t][2^basebit][n+1] int32"""
    N = len(u) - 1
    out = np.zeros(ksk.shape[-1], dtype=np.int64)
    out[-1] = u[N]
    prec = 1 << (32 - (1 + basebit * t_digits))
    for i in range(N):
        a = (int(u[i]) + prec) & M32
        for j in range(t_digits):
            d = (a >> (32 - basebit * (j + 1))) & ((1 << basebit) - 1)
            if d:
                out = (out - u32(ksk[i, j, d])) & M32
    return out


def keyswitch_row_sums(u, ksk, group=4, t_digits=8, basebit=2):
    """step 6 as the CUDA keyswitch evaluates it (csrc/peba1_cuda.cu: ksk_lut_build_kernel + keyswitch_lut_kernel): the
    `group` digits of a field select `group` key rows; their SUM is a table entry T[i][g][D], D = the field with the first
    digit in its top bits, and one entry per non-zero field is subtracted.  Integer addition mod 2^32 is associative, so
    this returns the words of keyswitch() above — the test says so on the golden chain."""
    N = len(u) - 1
    base = 1 << basebit
    out = np.zeros(ksk.shape[-1], dtype=np.int64)
    out[-1] = u[N]
    prec = 1 << (32 - (1 + basebit * t_digits))
    field = basebit * group

    def entry(i, g, D):  # what ksk_lut_build_kernel stores in row ((i * (t/group) + g) * (4^group - 1) + D - 1)
        s = np.zeros(ksk.shape[-1], dtype=np.int64)
        for m in range(group):
            d = (D >> (basebit * (group - 1 - m))) & (base - 1)
            if d:
                s = (s + u32(ksk[i, g * group + m, d])) & M32
        return s

    for i in range(N):
        a = (int(u[i]) + prec) & M32
        for g in range(t_digits // group):
            D = (a >> (32 - (g + 1) * field)) & ((1 << field) - 1)
            if D:
                out = (out - entry(i, g, D)) & M32
    return out


def gate(op, ca, cb, bk, ksk, l=3, Bgbit=7):
    return keyswitch(bootstrap_woks(prologue(op, ca, cb), bk, l, Bgbit), ksk)


def mux(cs, cx, cy, bk, ksk, l=3, Bgbit=7):
    """step 7 (bootsMUX, src/Math.cpp:261,277): two bootstraps without keyswitch, one keyswitch of their sum + 1/8"""
    t1 = (u32(cs) + u32(cx)) & M32
    t1[-1] = (t1[-1] - EIGHTH) & M32
    t2 = (u32(cy) - u32(cs)) & M32
    t2[-1] = (t2[-1] - EIGHTH) & M32
    u = (bootstrap_woks(t1, bk, l, Bgbit) + bootstrap_woks(t2, bk, l, Bgbit)) & M32
    u[-1] = (u[-1] + EIGHTH) & M32
    return keyswitch(u, ksk)


def as_i32(x):
    return (np.asarray(x, dtype=np.int64) & M32).astype(np.uint32).view(np.int32)
