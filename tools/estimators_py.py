"""Independent pure-Python restatement of the estimators, written from the published algorithms as recorded in
SURVEY.md Appendix A (A.3 HLL++ len, A.4 UltraLogLog FGRA / ML incl. small- and large-range corrections, A.5 HyperMinHash
cardinality / similarity) -- NOT from oracle/lash_oracle.c.  It exists to generate tests/golden/estimators.json
(tools/make_golden.py), which pins the oracle's estimator arithmetic against a second implementation in another language.
Python floats are IEEE doubles and math.pow/log/sqrt are the platform libm, so agreement is expected to a few ulp.
Test infrastructure only.
"""
from __future__ import annotations

import math

# ------------------------------------------------------------------------------------------------- A.3 HyperLogLog++
HLL_THRESH = {4: 10, 5: 20, 6: 40, 7: 80, 8: 220, 9: 400, 10: 900, 11: 1800, 12: 3100, 13: 6500, 14: 11500, 15: 20000,
              16: 50000, 17: 120000, 18: 350000}


def hll_alpha(p: int) -> float:
    return {4: 0.673, 5: 0.697, 6: 0.709}.get(p, 0.7213 / (1.0 + 1.079 / (1 << p)))


def hll_len(regs, p: int):
    """Returns (estimate, flagged): flagged = the estimate would need Google's empirical bias tables (not available)."""
    m = float(1 << p)
    zero = sum(1 for r in regs if r == 0)
    s = 0.0
    for r in regs:
        s += 2.0 ** -r
    if zero > 0:
        h = m * math.log(m / zero)
        if h <= HLL_THRESH[p]:
            return h, False
    e = hll_alpha(p) * m * m / s
    if e <= 5.0 * m:
        return float("nan"), True
    return e, False


# ------------------------------------------------------------------------------------------------- A.4 UltraLogLog
TAU = 0.8194911375910897
V = 0.6118931496978437
ETA = (4.663135422063788, 2.1378502137958524, 2.781144650979996, 0.9824082545153715)
ETA_X = ETA[0] - ETA[1] - ETA[2] + ETA[3]
ETA23X = (ETA[2] - ETA[3]) / ETA_X
ETA13X = (ETA[1] - ETA[3]) / ETA_X
ETA3012XX = (ETA[3] * ETA[0] - ETA[1] * ETA[2]) / (ETA_X * ETA_X)
P2T, P2MT, P4MT = 2.0 ** TAU, 2.0 ** -TAU, 4.0 ** -TAU
PHI_1 = ETA[0] / (P2T * (2.0 * P2T - 1.0))
P_INITIAL = ETA_X * (P4MT / (2.0 - P2MT))
INV_SQRT_FISHER = 0.7608621002725182
ML_BIAS = 0.48147376527720065


def reg_contribution(i: int) -> float:
    return ETA[i & 3] * 2.0 ** (-TAU * (3 + (i >> 2)))


def factor(p: int) -> float:
    m = float(1 << p)
    return m * m ** (1.0 / TAU) / (1.0 + V * (1.0 + TAU) / (2.0 * m))


def psi_prime(z: float, z2: float) -> float:
    return (z + ETA23X) * (z2 + ETA13X) + ETA3012XX


def sigma(z: float) -> float:
    if z <= 0.0:
        return ETA[3]
    if z >= 1.0:
        return math.inf
    pz, nz, s, pt = z, z * z, 0.0, ETA_X
    while True:
        old = s
        nn = nz * nz
        s += pt * (pz - nz) * psi_prime(nz, nn)
        if not s > old:
            return s / z
        pz, nz = nz, nn
        pt *= P2T


def phi(z: float, z2: float) -> float:
    if z <= 0.0:
        return 0.0
    if z >= 1.0:
        return PHI_1
    prev, pz = z2, z
    nz = math.sqrt(pz)
    t = P_INITIAL / (1.0 + nz)
    ps = psi_prime(pz, prev)
    s = nz * (ps + ps) * t
    while True:
        prev, pz = pz, nz
        old = s
        nz = math.sqrt(pz)
        nps = psi_prime(pz, prev)
        t *= P2MT / (1.0 + nz)
        s += nz * ((nps + nps) - (pz + nz) * ps) * t
        if not s > old:
            return s
        ps = nps


def ull_fgra(regs, p: int) -> float:
    m = 1 << p
    off = 4 * p + 4
    s = 0.0
    c0 = c4 = c8 = c10 = 0
    w = [0, 0, 0, 0]
    for r in regs:
        r2 = r - off
        if r2 < 0:
            if r2 < -8:
                c0 += 1
            if r2 == -8:
                c4 += 1
            if r2 == -4:
                c8 += 1
            if r2 == -2:
                c10 += 1
        elif r < 252:
            s += reg_contribution(r2)
        else:
            w[r - 252] += 1
    if c0 or c4 or c8 or c10:
        alpha = float(m + 3 * (c0 + c4 + c8 + c10))
        beta = float(m - c0 - c4)
        gamma = float(4 * c0 + 2 * c4 + 3 * c8 + c10)
        q = (math.sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha)
        z = (q * q) * (q * q)
        if c0:
            s += c0 * sigma(z)
        if c4:
            s += c4 * (P2MT * ETA_X) * psi_prime(z, z * z)
        if c8:
            s += c8 * (z * (P4MT * (ETA[0] - ETA[1])) + P4MT * ETA[1])
        if c10:
            s += c10 * (z * (P4MT * (ETA[2] - ETA[3])) + P4MT * ETA[3])
    if any(w):
        c = sum(w)
        alpha = float(m + 3 * c)
        beta = float(w[0] + w[1] + 2 * (w[2] + w[3]))
        gamma = float(m + 2 * w[0] + w[2] - w[3])
        z = math.sqrt((math.sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha))
        rz = math.sqrt(z)
        t = phi(rz, z) * c
        t += z * (1.0 + rz) * (w[0] * ETA[0] + w[1] * ETA[1] + w[2] * ETA[2] + w[3] * ETA[3])
        t += rz * ((w[0] + w[1]) * (z * (P2MT * (ETA[0] - ETA[2])) + P2MT * ETA[2]) +
                   (w[2] + w[3]) * (z * (P2MT * (ETA[1] - ETA[3])) + P2MT * ETA[3]))
        s += t * P2MT ** float(65 - p) / ((1.0 + rz) * (1.0 + z))
    return factor(p) * s ** (-1.0 / TAU)


def ml_contribute(r: int, b: list, p: int) -> int:
    r2 = r - 4 * p - 4
    if r2 < 0:
        ret = 4
        if r2 in (-2, -8):
            b[0] += 1
            ret -= 2
        if r2 in (-2, -4):
            b[1] += 1
            ret -= 1
        return (ret << (62 - p)) & ((1 << 64) - 1)
    k = r2 >> 2
    y0, y1 = r & 1, (r >> 1) & 1
    ret = 0xE000000000000000 - (y0 << 63) - (y1 << 62)
    b[k] += y0
    b[k + 1] += y1
    b[k + 2] += 1
    return ret >> (k + p)


def solve_ml(a: float, b: list, n: int, eps: float) -> float:
    if a == 0.0:
        return math.inf
    kmax = n
    while kmax >= 0 and b[kmax] == 0:
        kmax -= 1
    if kmax < 0:
        return 0.0
    kmin = kmax
    s1 = b[kmax]
    s2 = math.ldexp(float(b[kmax]), kmax)
    for k in range(kmax - 1, -1, -1):
        if b[k] > 0:
            s1 += b[k]
            s2 += math.ldexp(float(b[k]), k)
            kmin = k
    if s2 <= 1.5 * a:
        x = s1 / (0.5 * s2 + a)
    else:
        x = math.log1p(s2 / a) * (s1 / s2)
    dx, g_prev = x, 0.0
    while dx > x * eps:
        kappa = math.frexp(x)[1] - 1 + 2                      # ilogb(x) + 2
        xp = math.ldexp(x, -(max(kmax, kappa) + 1))
        xp2 = xp * xp
        h = xp - xp2 / 3.0 + (xp2 * xp2) * (1.0 / 45.0 - xp2 / 472.5)
        for _ in range(kappa - 1, kmax - 1, -1):
            hp = 1.0 - h
            h = (xp + h * hp) / (xp + hp)
            xp += xp
        g = b[kmax] * h
        for k in range(kmax - 1, kmin - 1, -1):
            hp = 1.0 - h
            h = (xp + h * hp) / (xp + hp)
            xp += xp
            g += b[k] * h
        g += x * a
        if g_prev < g <= s1:
            dx *= (g - s1) / (g_prev - g)
        else:
            dx = 0.0
        x += dx
        g_prev = g
    return x


def ull_ml(regs, p: int) -> float:
    m = 1 << p
    b = [0] * 66
    S = 0
    for r in regs:
        S = (S + ml_contribute(r, b, p)) & ((1 << 64) - 1)
    if S == 0:
        return 0.0 if regs[0] == 0 else math.inf
    b[63 - p] += b[64 - p]
    fac = float(2 * m)
    a = float(S) * fac * 2.0 ** -64
    eps = 1e-3 * INV_SQRT_FISHER / math.sqrt(float(m))
    return fac * solve_ml(a, b, 63 - p, eps) / (1.0 + ML_BIAS / m)


def ull_unpack(r: int) -> int:
    """(4 | (r & 3)) << ((r >> 2) - 2); valid non-empty registers are >= 4p-4 >= 8, so the shift is never negative."""
    return ((4 | (r & 3)) << ((r >> 2) - 2)) & ((1 << 64) - 1) if r else 0


def ull_pack(hp: int) -> int:
    nlz = 64 - hp.bit_length()
    return (((-(nlz + 1)) << 2) & 0xFF) | (((hp << (nlz + 1)) & ((1 << 64) - 1)) >> 62)


def ull_merge(a, b):
    out = []
    for x, y in zip(a, b):
        hp = ull_unpack(x) | ull_unpack(y)
        out.append(ull_pack(hp) if hp else 0)
    return out


# ------------------------------------------------------------------------------------------------- A.5 HyperMinHash
HMH_P, HMH_M, HMH_Q, HMH_R = 14, 16384, 6, 10
HMH_ALPHA = 0.7213 / (1.0 + 1.079 / HMH_M)
HMH_C = 0.169919487159739093975315012348


def hmh_beta(ez: float) -> float:
    zl = math.log(ez + 1.0)
    return (-0.370393911 * ez + 0.070471823 * zl + 0.17393686 * zl ** 2 + 0.16339839 * zl ** 3 - 0.09237745 * zl ** 4 +
            0.03738027 * zl ** 5 - 0.005384159 * zl ** 6 + 0.00042419 * zl ** 7)


def hmh_cardinality(regs) -> float:
    s, ez = 0.0, 0.0
    for r in regs:
        lz = r >> 10
        if lz == 0:
            ez += 1.0
        s += 2.0 ** -lz
    return HMH_ALPHA * HMH_M * (HMH_M - ez) / (hmh_beta(ez) + s)


def hmh_expected_collisions(n: float, m: float) -> float:
    if n < m:
        n, m = m, n
    if n > 2.0 ** 74:
        return float((1 << 64) - 1)
    if n > 2.0 ** (HMH_P + 5):
        d = (4.0 * n / m) / ((1.0 + n) / m) ** 2
        return HMH_C * 2.0 ** (HMH_P - HMH_R) * d + 0.5
    x = 0.0
    for i in range(1, 65):
        for j in range(1, 1025):
            if i != 64:
                den = 2.0 ** (HMH_P + HMH_R + i)
                b1, b2 = (1024.0 + j) / den, (1025.0 + j) / den
            else:
                den = 2.0 ** (HMH_P + HMH_R + i - 1)
                b1, b2 = j / den, (j + 1.0) / den
            x += ((1.0 - b2) ** n - (1.0 - b1) ** n) * ((1.0 - b2) ** m - (1.0 - b1) ** m)
    return (x * HMH_P + 0.5) / HMH_P


def hmh_similarity(a, b) -> float:
    c = sum(1 for x, y in zip(a, b) if x != 0 and x == y)
    n = sum(1 for x, y in zip(a, b) if x != 0 or y != 0)
    if c == 0:
        return 0.0
    ec = hmh_expected_collisions(hmh_cardinality(a), hmh_cardinality(b))
    if c < ec:
        return 0.0
    return (c - ec) / n


# ------------------------------------------------------------------------------------------------- distance (main.rs:415-423)
def frac_from(card_a: float, card_b: float, union: float, hll: bool) -> float:
    sim = (card_a + card_b - union) / union
    if hll:
        s = 0.0 if math.isnan(sim) else max(sim, 0.0)      # f64::max: NaN -> 0 (utils.rs:362)
    else:
        s = 0.0 if sim < 0.0 else sim                      # utils.rs:274: NaN propagates
    return 2.0 * s / (1.0 + s)


def mash(frac: float, k: int, model: int) -> float:
    if model == 1:
        return min(-math.log(frac) / k, 1.0) if frac > 0.0 else 1.0
    return 1.0 - frac ** (1.0 / k)
