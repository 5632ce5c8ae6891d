"""Synthetic Theil-Sen inputs with 0/0 slopes (TEST INFRASTRUCTURE).  A case = cleaned (signal, k-mer rank) vectors as
normaliseEvents builds them plus rough scalings; `dup` pairs of sampled points are made identical (same signal, same
rank), each giving one NaN among the 499 500 slopes the reference std::sorts (event_handling.cpp:67-78)."""
import numpy as np


def make_case(seed: int, model_mean: np.ndarray, n: int = 1300, dup: int = 1, noise: float = 2.0):
    rng = np.random.default_rng(seed)
    ranks = rng.integers(0, model_mean.size, n).astype(np.uint32)
    shift, scale = 90.0 + rng.normal(0, 3), 13.0 + rng.normal(0, 0.5)
    sig = (model_mean[ranks] - 95.0) / 14.0 * scale * (1.0 + rng.normal(0, 0.01)) + shift + rng.normal(0, noise, n)
    eff = n - 100
    skip = eff // 1000 if eff > 1000 else 1
    npnt = min(eff, 1000)
    for _ in range(dup):
        a, b = rng.choice(npnt, 2, replace=False)
        ia, ib = 50 + int(a) * skip, 50 + int(b) * skip
        sig[ib] = sig[ia]
        ranks[ib] = ranks[ia]
    return sig.astype(np.float64), ranks, float(shift), float(scale)


def cases(model_mean: np.ndarray, count: int = 48):
    """count cases: mostly one NaN; every 8th has none, every 12th has two (not emulated on the device: the run must
    agree with the reference there too or the test says so), a few longer reads with skip > 1."""
    out = []
    for k in range(count):
        dup = 0 if k % 8 == 7 else (2 if k % 12 == 11 else 1)
        n = 1300 if k % 5 else 1100 + 531 * (k % 7)
        out.append(make_case(1000 + k, model_mean, n=n, dup=dup))
    return out
