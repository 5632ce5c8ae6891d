"""CPU model of csrc/seg_scan.cu (one warp = 32 lanes, chunks of 32 blocks x 64 samples): the same phases, the same
float64 operations in the same order, so that the kernel's algebra -- magic-constant rounding of the addends, the
two-state transducer per block, the warp scan of transducers, the acceptance test and the literal fallback -- can be
checked against scrappie's sequential loop (/root/reference/src/scrappie/event_detection.c:35-48) without a GPU.
Python floats are IEEE doubles and every operation below is a single rounding, like the kernel's __dadd_rn/__dmul_rn."""
import math

import numpy as np

BLOCK = 64
TWO52 = 4503599627370496.0
NO_MAP = None


def binade_of(s: float) -> int:
    return math.frexp(s)[1] - 1


def fold_sq(a: float, inv_u: float, d0: float, d1: float):
    A = a * inv_u
    if not (A < TWO52):
        return None
    g0 = (A + TWO52) - TWO52
    diff = g0 - A
    if abs(diff) == 0.5:
        g1 = g0 - (diff + diff)
        odd0 = int(d0) & 1
        odd1 = int(d1) & 1
        return d0 + (g1 if odd0 else g0), d1 + (g0 if odd1 else g1)
    return d0 + g0, d1 + g0


def sequential(x: np.ndarray):
    """scrappie: prefix sums with a rounding at every step; returns the values at every 64th sample and the totals."""
    s = q = 0.0
    cs, cq = [], []
    for j, xv in enumerate(x.astype(np.float64)):
        if j % BLOCK == 0:
            cs.append(s); cq.append(q)
        s = s + xv
        q = q + xv * xv
    return np.array(cs), np.array(cq), s, q


def scan(x: np.ndarray, stats: dict | None = None):
    """The kernel: returns (ck_sum, ck_sq, tot_sum, final sumsq, exact_flag)."""
    x = x.astype(np.float64)
    N = x.size
    nb = (N + BLOCK - 1) // BLOCK
    cs, cq = np.zeros(nb), np.zeros(nb)
    s_act = q_act = 0.0
    abs_sum, min_exp = 0.0, None
    fast = slow = literal = 0
    for b0 in range(0, nb, 32):
        lanes = range(min(32, nb - b0))
        S, Q = [0.0] * 32, [0.0] * 32
        for L in lanes:
            for xv in x[(b0 + L) * BLOCK:(b0 + L + 1) * BLOCK]:
                S[L] = S[L] + xv
                abs_sum = abs_sum + abs(xv)
                Q[L] = Q[L] + xv * xv
                if xv != 0.0:
                    e = math.frexp(float(np.float32(xv)))[1] - 1
                    min_exp = e if min_exp is None else min(min_exp, e)
        inS, inQ = list(S), list(Q)
        d = 1
        while d < 32:                                      # Hillis-Steele, as the shuffles do it
            nS, nQ = list(inS), list(inQ)
            for L in range(d, 32):
                nS[L] = inS[L] + inS[L - d]; nQ[L] = inQ[L] + inQ[L - d]
            inS, inQ = nS, nQ
            d <<= 1
        eq, m0, m1 = [NO_MAP] * 32, [0] * 32, [0] * 32
        for L in lanes:
            cs[b0 + L] = s_act + (inS[L] - S[L])
            pq = q_act + (inQ[L] - Q[L])
            if pq > 0.0:
                e = binade_of(pq)
                inv_u = math.ldexp(1.0, 52 - e)
                d0 = d1 = 0.0
                ok = True
                for xv in x[(b0 + L) * BLOCK:(b0 + L + 1) * BLOCK]:
                    r = fold_sq(xv * xv, inv_u, d0, d1)
                    if r is None:
                        ok = False
                        break
                    d0, d1 = r
                if ok:
                    eq[L], m0[L], m1[L] = e, int(d0), int(d1)
        s_act = s_act + inS[31]
        e_act = binade_of(q_act) if q_act > 0.0 else NO_MAP
        done = False
        if e_act is not NO_MAP and all(eq[L] == e_act for L in lanes):
            S0 = int(q_act * math.ldexp(1.0, 52 - e_act))
            f0, f1 = list(m0), list(m1)
            d = 1
            while d < 32:
                n0, n1 = list(f0), list(f1)
                for L in range(d, 32):
                    a0, a1 = f0[L - d], f1[L - d]
                    n0[L] = a0 + (f1[L] if a0 & 1 else f0[L])
                    n1[L] = a1 + (f0[L] if a1 & 1 else f1[L])
                f0, f1 = n0, n1
                d <<= 1
            x0 = [0] + f0[:31]; x1 = [0] + f1[:31]
            Sb = [S0 + (x1[L] if S0 & 1 else x0[L]) for L in range(32)]
            Se = [S0 + (f1[L] if S0 & 1 else f0[L]) for L in range(32)]
            if all(m0[L] < (1 << 52) and m1[L] < (1 << 52) and Se[L] < (1 << 53) for L in lanes):
                u = math.ldexp(1.0, e_act - 52)
                for L in lanes:
                    cq[b0 + L] = float(Sb[L]) * u
                q_act = float(Se[len(lanes) - 1]) * u
                done = True
                fast += 1
        if not done:
            slow += 1
            for L in lanes:
                cq[b0 + L] = q_act
                applied = False
                if eq[L] is not NO_MAP and m0[L] < (1 << 52) and m1[L] < (1 << 52) and q_act > 0.0 and binade_of(q_act) == eq[L]:
                    Sx = int(q_act * math.ldexp(1.0, 52 - eq[L]))
                    Sx += m1[L] if Sx & 1 else m0[L]
                    if Sx < (1 << 53):
                        q_act = float(Sx) * math.ldexp(1.0, eq[L] - 52)
                        applied = True
                if not applied:
                    literal += 1
                    for xv in x[(b0 + L) * BLOCK:(b0 + L + 1) * BLOCK]:
                        q_act = q_act + xv * xv
    exact = True
    if min_exp is not None:
        exact = abs_sum * 1.0001 < math.ldexp(1.0, min_exp - 23 + 53)
    if stats is not None:
        stats.update(fast_chunks=fast, serial_chunks=slow, literal_blocks=literal, blocks=nb)
    return cs, cq, s_act, q_act, exact
