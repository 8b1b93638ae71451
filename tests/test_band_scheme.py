"""The banded-staircase scheme of ed_band_kernel / seg_band_kernel (delly_b200/csrc/myers.cuh, edit_distance.cu, edit_path.cu) as an executable
specification on the CPU: a lane-by-lane simulation of the device schedule (G lanes per job, 64-row blocks, one-step skew, block hand-over after every
64 columns, certified range from the staircase geometry) against a plain full-matrix DP. It pins the three claims the kernels rest on:
  1. a value the pass certifies (value <= kvalid, or the staircase covers the matrix) is the exact NW distance, and an uncertified value never
     under-estimates it;
  2. the traceback from the stored per-column (Pv, Ph) bits — up iff Pv, else left iff Ph, else diagonal — is, op for op, the full-matrix traceback
     with the reference's priority (src/edlib.cpp:1021-1131);
  3. Hirschberg's split row found from two banded half columns run in the PARENT's band equals the one found from full half columns
     (src/edlib.cpp:1303-1338: interior rows ascending, then 0, then m).
The CUDA kernels themselves are checked against the reference in tests/test_edit_distance.py and tests/test_edit_path.py (-m gpu)."""
import numpy as np

M64 = (1 << 64) - 1


def block64(Pv, Mv, Eq, hin):   # Myers / Hyyro block update as myers.cuh: block64
    hin_neg = 1 if hin < 0 else 0
    Xv = Eq | Mv
    Eq |= hin_neg
    Xh = ((((Eq & Pv) + Pv) & M64) ^ Pv) | Eq
    Ph = Mv | (~(Xh | Pv) & M64)
    Mh = Pv & Xh
    ph_out = Ph
    hout = (Ph >> 63) - (Mh >> 63)
    Ph = (Ph << 1) & M64
    Mh = (Mh << 1) & M64
    Mh |= hin_neg
    Ph |= ((hin + 1) & 0xffffffff) >> 1
    Pv = Mh | (~(Xv | Ph) & M64)
    Mv = Ph & Xv
    return Pv, Mv, hout, ph_out


def band_plan(G, m, n):   # myers.cuh: band_plan
    d = n - m
    up, lo = max(d, 0), max(-d, 0)
    best, A = -10 ** 9, 0
    for a in range(G):
        h = min(64 * a - up, 64 * (G - 1 - a) - lo)
        if h > best:
            best, A = h, a
    kvalid = 2 * best + up + lo if best >= 0 else -1
    whole = ((n - 1) >> 6) <= A and 64 * (G - A) >= m
    return A, kvalid, whole


def band_pass(q, t, G, pm, pn, trace):
    """one pass of G lanes over (q, t) in the band of the problem pm x pn; returns (corner or None, last column, ops or None)"""
    m, n = len(q), len(t)
    nblk = (m + 63) >> 6
    A, _, _ = band_plan(G, pm, pn)
    peq = {}
    for b in range(nblk):
        for sy in set(q) | set(t):
            v = 0
            for i in range(64):
                if b * 64 + i < m and q[b * 64 + i] == sy:
                    v |= 1 << i
            peq[(b, sy)] = v
    Jend = (n - 1) >> 6
    steps = 65 * Jend + ((n - 1) & 63) + G
    Pv, Mv = [M64] * G, [0] * G
    blk, J, c = [p - A for p in range(G)], [0] * G, [-p for p in range(G)]
    sc, hout = [64 * (blk[p] + 1) for p in range(G)], [1] * G
    store = {}
    for st in range(steps):
        hin_sh = [hout[p - 1] if p > 0 else 0 for p in range(G)]
        for p in range(G):
            if 0 <= c[p] < 64:
                col = (J[p] << 6) + c[p]
                if col < n and 0 <= blk[p] < nblk:
                    hin = 1 if (p == 0 or blk[p] == 0) else hin_sh[p]
                    Pv[p], Mv[p], hout[p], Ph = block64(Pv[p], Mv[p], peq[(blk[p], t[col])], hin)
                    sc[p] += hout[p]
                    store[st * G + p] = (Pv[p], Ph)
        slot = [(Pv[p], Mv[p], sc[p]) for p in range(G)]
        for p in range(G):
            if c[p] == 64 and ((J[p] + 1) << 6) < n:
                if p < G - 1:
                    assert c[p + 1] == 63          # the lane below has just finished the chunk
                    Pv[p], Mv[p], sc[p] = slot[p + 1]
                else:
                    Pv[p], Mv[p], sc[p] = M64, 0, sc[p] + 64
                blk[p] += 1; J[p] += 1; c[p] = -1
        for p in range(G):
            c[p] += 1
    col = [None] * (m + 1)
    col[0] = n
    for p in range(G):
        if 0 <= blk[p] < nblk:
            v = sc[p]
            for j in range(63, -1, -1):
                i = 64 * blk[p] + j + 1
                if i <= m:
                    col[i] = v
                v -= ((Pv[p] >> j) & 1) - ((Mv[p] >> j) & 1)
    plast = (nblk - 1) - (Jend - A)
    corner = col[m] if 0 <= plast < G else None
    if not trace or corner is None:
        return corner, col, None
    r, cc, ops = m - 1, n - 1, []
    while r >= 0 and cc >= 0:
        b, Jc = r >> 6, cc >> 6
        pp = b - (Jc - A)
        assert 0 <= pp < G
        pv, ph = store[(65 * Jc + (cc & 63) + pp) * G + pp]
        bit = r & 63
        if (pv >> bit) & 1:
            ops.append(1); r -= 1
        elif (ph >> bit) & 1:
            ops.append(2); cc -= 1
        else:
            ops.append(0 if q[r] == t[cc] else 3); r -= 1; cc -= 1
    ops += [1] * (r + 1) + [2] * (cc + 1)
    return corner, col, ops[::-1]


def full_dp(q, t):
    m, n = len(q), len(t)
    D = np.zeros((m + 1, n + 1), np.int32)
    D[:, 0] = np.arange(m + 1)
    D[0, :] = np.arange(n + 1)
    tt = np.frombuffer(t, np.uint8)
    for i in range(1, m + 1):
        best = np.minimum(D[i - 1, :-1] + (tt != q[i - 1]), D[i - 1, 1:] + 1)
        row = D[i]
        for j in range(1, n + 1):
            row[j] = min(best[j - 1], row[j - 1] + 1)
    r, c, ops = m, n, []
    while r > 0 or c > 0:
        if r == 0: ops.append(2); c -= 1
        elif c == 0: ops.append(1); r -= 1
        elif D[r, c] == D[r - 1, c] + 1: ops.append(1); r -= 1
        elif D[r, c] == D[r, c - 1] + 1: ops.append(2); c -= 1
        else: ops.append(0 if q[r - 1] == t[c - 1] else 3); r -= 1; c -= 1
    return D, ops[::-1]


def _pairs(seed, count):
    rng = np.random.default_rng(seed)
    AL = b"ACGT"
    for it in range(count):
        L = int(rng.integers(65, 330))
        t = bytes(rng.choice(list(AL), size=L).tolist())
        kind = it % 3
        q = bytearray(t)
        if kind == 0:
            for p in rng.integers(0, L, size=int(L * rng.choice([0.0, 0.03, 0.08]))):
                q[p] = AL[rng.integers(0, 4)]
        elif kind == 1:
            rate, out = float(rng.choice([0.03, 0.12, 0.2])), bytearray()
            for ch in q:
                x = rng.random()
                if x < rate / 3: continue
                if x < 2 * rate / 3: out.append(AL[rng.integers(0, 4)]); continue
                out.append(ch)
                if x < rate: out.append(AL[rng.integers(0, 4)])
            q = out
        else:
            g = int(rng.integers(1, 120)); a = int(rng.integers(0, max(1, L - g)))
            q = q[:a] + q[a + g:]
        q = bytes(q)
        if len(q) < 2: continue
        yield (t, q) if rng.random() < 0.5 else (q, t)


def test_certified_values_are_exact_and_tracebacks_and_split_rows_match_the_full_matrix():
    certified = traced = splits = 0
    for q, t in _pairs(7, 36):
        D, ops = full_dp(q, t)
        d, m, n = int(D[len(q), len(t)]), len(q), len(t)
        for G in (2, 4, 8):
            A, kvalid, whole = band_plan(G, m, n)
            corner, _, bops = band_pass(q, t, G, m, n, trace=(kvalid >= d))   # the host only sends a leaf to a class that certifies its optimum
            if corner is None:
                continue
            if whole or corner <= kvalid:
                certified += 1
                assert corner == d, (m, n, G, corner, d, kvalid)
            else:
                assert corner >= d, (m, n, G, corner, d)
            if kvalid >= d:
                traced += 1
                assert bops == ops, (m, n, G, d)
                # Hirschberg: both halves in the parent's band
                h = n // 2
                if h == 0 or h == n:
                    continue
                _, fcol, _ = band_pass(q, t[:h], G, m, n, trace=False)
                _, bcol, _ = band_pass(q[::-1], t[h:][::-1], G, m, n, trace=False)
                Df, _ = full_dp(q, t[:h])
                Db, _ = full_dp(q[::-1], t[h:][::-1])
                INF = 1 << 28
                def first_split(F, B):
                    for i in list(range(1, m)) + [0, m]:
                        if F(i) + B(m - i) == d:
                            return i
                    return -1
                want = first_split(lambda i: int(Df[i, h]), lambda i: int(Db[i, n - h]))
                got = first_split(lambda i: fcol[i] if fcol[i] is not None else INF, lambda i: bcol[i] if bcol[i] is not None else INF)
                assert want == got and want >= 0, (m, n, G, want, got)
                splits += 1
                for i in range(m + 1):   # banded columns never under-estimate
                    assert fcol[i] is None or fcol[i] >= Df[i, h]
    assert certified >= 40 and traced >= 40 and splits >= 30, (certified, traced, splits)
