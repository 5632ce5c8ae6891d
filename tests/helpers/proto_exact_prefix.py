"""Prototype (CPU, no product code): the sequentially ROUNDED prefix sums of scrappie's compute_sum_sumsq
(src/scrappie/event_detection.c:35-48) computed WITHOUT a per-sample serial chain, bit for bit.

Why.  sumsq[i+1] = fl(sumsq[i] + x[i]*x[i]) rounds at every step, so the value depends on the summation order and the
checkpoint kernel of csrc/seg.cu walks each read serially (one lane per read, ~31 cycles per sample): a pure latency
chain that dominates segmentation on short batches.

Idea.  All addends a_i = x_i^2 are >= 0 and exactly representable, so s is non-decreasing and visits each binade once.
Inside a binade with ulp u, write s = S*u (S integer) and a = q*u + r (0 <= r < u).  Then
    fl(s + a) = (S + q + c) * u,   c = 0 if r < u/2,  1 if r > u/2,  (S + q) mod 2 if r == u/2   (ties to even)
i.e. the step is S -> S + d[S mod 2] with a PAIR of integers (d[0], d[1]) that depends on a and u only.  Such maps compose
associatively:  (g o f).d[p] = f.d[p] + g.d[p xor (f.d[p] & 1)].  A tile of samples therefore collapses to one pair by a
parallel reduction (integer adds and selects), the per-read serial chain shrinks from one step per SAMPLE to one step per
TILE, and the tiles whose range of possible values straddles a power of two (at most one or two per binade, ~40 per
read) are the only ones that still need the literal sample-by-sample loop.

This script checks the claim against the sequential double-precision loop on realistic and adversarial inputs.
Run: python tests/helpers/proto_exact_prefix.py
"""
import numpy as np

TILE = 512


def seq_prefix(a):
    """sum[i+1] = fl(sum[i] + a[i]) exactly as the C loop does it."""
    out = np.empty(a.size + 1)
    s = 0.0
    out[0] = 0.0
    for i, v in enumerate(a.tolist()):
        s = s + v
        out[i + 1] = s
    return out


def ulp_exp(s):
    """exponent e with s in [2^e, 2^(e+1)); ulp = 2^(e-52)"""
    m, e = np.frexp(s)
    return int(e) - 1


def tile_maps(a, e):
    """(d0, d1) of every element of `a` for running sums in binade e (ulp u = 2^(e-52)); exact integer arithmetic."""
    sh = e - 52
    # a = q*u + r exactly: scale by 2^-sh (a power of two: exact unless it underflows, which these magnitudes do not)
    scaled = np.ldexp(a, -sh)
    q = np.floor(scaled)
    r = scaled - q                                   # exact: both are multiples of 2^-k with few bits
    qi = q.astype(np.int64)
    up = (r > 0.5).astype(np.int64)
    tie = r == 0.5
    d0 = qi + up
    d1 = qi + up
    odd = (qi & 1) == 1
    d0 = np.where(tie, np.where(odd, qi + 1, qi), d0)
    d1 = np.where(tie, np.where(odd, qi, qi + 1), d1)
    return d0, d1


def compose(f0, f1, g0, g1):
    """(g o f): apply f, then g"""
    h0 = f0 + np.where((f0 & 1) == 1, g1, g0)
    h1 = f1 + np.where((f1 & 1) == 0, g1, g0)       # start odd: parity after f is 1 xor (f1 & 1)
    return h0, h1


def reduce_tile(d0, d1):
    """pairwise (tree) reduction, as a warp/CTA would do it"""
    while d0.size > 1:
        if d0.size & 1:
            d0 = np.append(d0, 0)
            d1 = np.append(d1, 0)                    # identity map
        d0, d1 = compose(d0[0::2], d1[0::2], d0[1::2], d1[1::2])
    return int(d0[0]), int(d1[0])


def tiled_prefix_at_tile_ends(a, stats):
    """Values of the rounded running sum at every tile boundary, with one serial step per tile wherever the tile lies inside
    one binade, and the literal loop otherwise."""
    n_tiles = (a.size + TILE - 1) // TILE
    ends = np.empty(n_tiles)
    s = 0.0
    for t in range(n_tiles):
        seg = a[t * TILE:(t + 1) * TILE]
        exact_upper = s + float(seg.sum()) * (1 + 1e-12) + 1e-300          # a cheap, safe bound on where the tile can end
        if s > 0.0 and ulp_exp(s) == ulp_exp(exact_upper) and ulp_exp(s) - 52 > -1000:
            e = ulp_exp(s)
            d0, d1 = reduce_tile(*tile_maps(seg, e))
            S = int(np.ldexp(s, 52 - e))                                  # exact: s is a multiple of its ulp
            S += d1 if (S & 1) else d0
            assert S < (1 << 53), "left the binade although the bound said it would not"
            s = float(np.ldexp(float(S), e - 52))
            stats["parallel"] += 1
        else:
            for v in seg.tolist():
                s = s + v
            stats["serial"] += 1
        ends[t] = s
    return ends


def check(name, x):
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    for label, a in (("sumsq", x * x), ("sum", np.abs(x))):
        ref = seq_prefix(a)
        stats = {"parallel": 0, "serial": 0}
        got = tiled_prefix_at_tile_ends(a, stats)
        idx = np.minimum(np.arange(1, got.size + 1) * TILE, a.size)
        assert np.array_equal(got, ref[idx]), (name, label)
        print(f"{name:28s} {label:6s} n={a.size:8d}  tiles: {stats['parallel']:6d} by composition, {stats['serial']:3d} serial"
              f"  -> bit-identical")


def main():
    rng = np.random.default_rng(1)
    # realistic: int16 DAC -> pA exactly as pod5.cpp:60, levels 60..130 pA
    dac = rng.integers(200, 900, 400_000).astype(np.int16)
    pa = (dac.astype(np.float32) + np.float32(10.0)) * np.float32(0.1755)
    check("pod5-like 400k samples", pa)
    # adversarial: many exact ties (values with few mantissa bits), tiny and huge magnitudes mixed, long plateaus
    check("few-bit values (ties)", rng.integers(1, 64, 300_000) * 0.5)
    check("wide dynamic range", np.exp(rng.normal(0, 4, 200_000)))
    check("plateaus", np.repeat(rng.integers(1, 2000, 400) * 0.125, 700))
    check("powers of two", np.ldexp(1.0, rng.integers(-8, 9, 200_000)))


if __name__ == "__main__":
    main()


# ---- the kernel-shaped version: 64-sample blocks, SIGNED addends, maps built under a PREDICTED binade --------------------
# (mirrors dnascent_b200/csrc/seg_scan.cu: phase A block sums, B approximate prefix -> predicted binade per block,
#  C per-block maps under that binade, D one serial step per block with an exact check and a literal fallback)
BLOCK = 64


def block_chain(a):
    """rounded running sums of the signed, exactly representable addends `a` at every block boundary"""
    nb = (a.size + BLOCK - 1) // BLOCK
    pad = np.zeros(nb * BLOCK)
    pad[:a.size] = a                                             # +0.0 addends change nothing
    blk = pad.reshape(nb, BLOCK)
    bs = blk.sum(axis=1)                                         # phase A (any summation order: only used to predict)
    ba = np.abs(blk).sum(axis=1)
    pre = np.concatenate([[0.0], np.cumsum(bs)[:-1]])            # phase B: approximate running sum at each block start
    e_pred = np.frexp(np.where(pre > 0, pre, 1.0))[1] - 1        # predicted binade (pre <= 0: no prediction)
    maps = []
    for b in range(nb):                                          # phase C (parallel over blocks on the device)
        if pre[b] <= 0:
            maps.append(None)
            continue
        e = int(e_pred[b])
        sh = e - 52
        scaled = np.ldexp(blk[b], -sh)
        q = np.floor(scaled)                                     # floor, also for negative addends: r in [0, 1)
        r = scaled - q
        if not np.all(np.isfinite(scaled)) or np.any(np.abs(q) >= 2.0 ** 53):
            maps.append(None)
            continue
        d0, d1 = 0, 0                                            # identity
        for qi, ri in zip(q.astype(np.int64).tolist(), r.tolist()):
            up = 1 if ri > 0.5 else 0
            if ri == 0.5:
                g0, g1 = (qi + 1, qi) if (qi & 1) else (qi, qi + 1)
            else:
                g0 = g1 = qi + up
            d0 = d0 + (g1 if (d0 & 1) else g0)                   # compose: this element after the block so far
            d1 = d1 + (g0 if (d1 & 1) else g1)
        maps.append((e, d0, d1))
    out = np.empty(nb + 1)
    s = 0.0
    stats = [0, 0]
    for b in range(nb):                                          # phase D: one lane per read
        out[b] = s
        m = maps[b]
        done = False
        if m is not None and s > 0:
            e, d0, d1 = m
            lo, hi = 2.0 ** e, 2.0 ** (e + 1)
            # exact precondition: the actual running sum is in the predicted binade and cannot leave it inside the block
            if lo <= s < hi and s - ba[b] * 1.0000001 >= lo and s + ba[b] * 1.0000001 < hi:
                S = int(np.ldexp(s, 52 - e))
                S += d1 if (S & 1) else d0
                s = float(np.ldexp(float(S), e - 52))
                done = True
        if not done:
            for v in blk[b].tolist():
                s = s + v
        stats[0 if done else 1] += 1
    out[nb] = s
    return out, stats


def check_blocks(name, x):
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    for label, a in (("sumsq", x * x), ("sum", x)):
        ref = seq_prefix(a)
        got, stats = block_chain(a)
        idx = np.minimum(np.arange(got.size) * BLOCK, a.size)
        assert np.array_equal(got, ref[idx]), (name, label)
        print(f"{name:28s} {label:6s} n={a.size:8d}  blocks: {stats[0]:6d} by composition, {stats[1]:5d} literal  -> bit-identical")


def main_blocks():
    rng = np.random.default_rng(2)
    dac = rng.integers(200, 900, 200_000).astype(np.int16)
    check_blocks("pod5-like", (dac.astype(np.float32) + np.float32(10.0)) * np.float32(0.1755))
    check_blocks("signed, zero-mean", rng.normal(0, 50, 100_000))
    check_blocks("signed with drift", rng.normal(3, 50, 100_000))
    check_blocks("few-bit signed (ties)", rng.integers(-64, 64, 100_000) * 0.5)
    check_blocks("wide dynamic range", np.exp(rng.normal(0, 4, 100_000)) * rng.choice([-1.0, 1.0], 100_000))


if __name__ == "__main__":
    main_blocks()
