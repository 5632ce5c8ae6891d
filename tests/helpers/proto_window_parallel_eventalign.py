"""Prototype (CPU, no product code): eventalign (src/alignment.cpp:547-744) with its windows computed in PARALLEL.

Why.  The reference walks a read's ~50-base windows serially: window w+1 starts one past the last match state of window
w (alignment.cpp:740-741).  The GPU kernel of this round keeps that chain (one warp per read), so a read's latency is
its length (34 ms per 10 kb, ~0.85 s for a 250-kb read) and the longest read bounds a submission.

Observation (measured with DNBO_EA_WINDOW_LOG on simulated reads): in ~99.7 % of the windows the Viterbi path ends with
a match in the LAST state on the LAST event, i.e. the next window starts at reference_index + (windowLength - 8) with
all of the window's events consumed.  In that case the next window is a function of the reference alone: its length
comes from the breakpoint rule at its start, and its events are exactly the aligned events whose k-mer lies in its
query range (readHead is then only a lower bound of the first such event and does not change the result).

Algorithm checked here.
  round 0   follow the "every window advances fully" chain (no Viterbi needed to know it) and run builtinViterbi on
            all of its windows independently -- this is the parallel part
  verify    walk the TRUE chain with the results at hand: a window is looked up by (reference_index, readHead-relevant
            flag); where the true chain leaves the speculative one, the windows it needs are collected, with the same
            optimistic continuation behind them
  round k   run the missing windows, verify again; stop when the walk finds everything it needs
The true chain usually REJOINS the speculative one a window later (the next breakpoint is the same absolute position),
so the extra rounds touch a fraction of a percent of the windows.

The script asserts that the records are identical to the oracle port's serial eventalign and prints how many windows
each round ran.  Run: python tests/helpers/proto_window_parallel_eventalign.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dnascent_b200 import synth          # noqa: E402
from oracle import portbind              # noqa: E402

K = 9


class Read:
    """What eventalign reads from a normalised DNAscent::read, plus the serial chain's records as the truth."""

    def __init__(self, P, mean, ref, r2q, al_e, al_k, evm, shift, scale, epb, window=50, serial=None):
        self.P, self.mean, self.W = P, mean, window
        self.ref = ref
        self.r2q = np.asarray(r2q, dtype=np.int64)
        self.al_e, self.al_k = np.asarray(al_e, dtype=np.int64), np.asarray(al_k, dtype=np.int64)
        self.evm = np.asarray(evm, dtype=np.float64)
        self.shift, self.scale, self.epb = shift, scale, epb
        self.kmean = mean[synth.kmer_ranks(self.ref)]                      # model mean of the k-mer at every position
        self.defined = np.frombuffer(self.ref, dtype=np.uint8)
        self.defined = np.isin(self.defined, np.frombuffer(b"ACGT", dtype=np.uint8))
        self.serial = serial if serial is not None else P.eventalign(
            self.ref, self.r2q.astype(np.int32), self.al_e, self.al_k, self.evm, self.shift, self.scale, self.epb, mean, window)

    @classmethod
    def from_synth(cls, P, sr, mean, window=50):
        p = P.normalise(sr.raw, sr.basecall, sr.refseq, sr.query_to_ref, mean)
        assert p["status"] == 0
        q2r = np.asarray(sr.query_to_ref)
        r2q = np.zeros(len(sr.refseq), dtype=np.int64)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]
        return cls(P, mean, sr.refseq, r2q, p["align_event"], p["align_kmer"], p["event_mean"], p["shift"], p["scale"],
                   p["events_per_base"], window)

    # ---- pieces of alignment.cpp:556-641 that do not need the Viterbi result ----
    def window_len(self, ri):
        """(windowLength, usable) at reference_index ri: breakpoint rule :565-595 and the referenceDefined checks"""
        rlen, W = len(self.ref), self.W
        to_end = rlen - ri
        wl = min(to_end, W)
        if to_end > 1.5 * W:
            bl = int(1.5 * wl)
            if not self.defined[ri:ri + bl].all():
                return wl, False
            m = self.kmean
            for i in range(wl, int(np.ceil(1.5 * wl - K - 1))):
                if abs(m[ri + i] - m[ri + i + 1]) > 0.75 and abs(m[ri + i] - m[ri + i - 1]) > 0.75:
                    wl = i + K
                    break
        return wl, bool(self.defined[ri:ri + wl].all())

    def gather(self, ri, wl, read_head):
        """events of the window (:611-632): (new readHead, observations, their event indices, lo, hi)"""
        lo, hi = self.r2q[ri], self.r2q[ri + wl - K + 1]
        obs, ev = [], []
        first = True
        for j in range(read_head, self.al_k.size):
            k = self.al_k[j]
            if lo <= k < hi:
                if first:
                    read_head, first = j, False
                em = self.evm[self.al_e[j]]
                if 0.0 < em < 250.0:
                    obs.append(em)
                    ev.append(int(self.al_e[j]))
            if k >= hi:
                break
        return read_head, np.asarray(obs), ev, int(lo), int(hi)

    def first_in_range(self, lo):
        return int(np.searchsorted(self.al_k, lo, side="left"))


def run_window(R, ri, wl, read_head):
    """builtinViterbi on one window + the two passes over its state labels (:655-741).  Returns (None, readHead) for a
    skipped window (fewer than two observations), else (records, last_m_ref, last_m_ev, readHead after the gather)."""
    rh, obs, ev, lo, hi = R.gather(ri, wl, read_head)
    if obs.size < 2:
        return None, rh                # skipped (:641) -- but the gather has already moved readHead (:617-620)
    _, idx, typ = R.P.builtin_viterbi(obs, R.ref[ri:ri + wl], R.shift, R.scale, R.epb, R.mean)
    indel = (hi - lo) - (wl - K + 1)
    last_m_ev = last_m_ref = 0
    e = 0
    for i, t in zip(idx, typ):
        if t == 1:
            last_m_ev, last_m_ref = e, int(i)
        if t != 0:
            e += 1
    recs = []
    e = 0
    for i, t in zip(idx, typ):
        if t == 0:
            continue
        if t == 1 or (t == 2 and e < last_m_ev):
            recs.append((ev[e], ri + int(i), int(t), indel))
        e += 1
    return recs, last_m_ref, last_m_ev, rh


def speculative(R):
    rlen = len(R.ref)
    cache = {}                      # (ri, read_head or None) -> run_window result; None key part = "readHead does not matter"
    rounds = []

    def key_for(ri, wl, read_head):
        # the window's events are independent of readHead iff readHead <= first aligned event of its query range
        lo = R.r2q[ri]
        return (ri, None) if read_head <= R.first_in_range(lo) else (ri, read_head)

    def walk(collect):
        """the true chain, as far as the cache allows; unknown windows are assumed to advance fully"""
        ri, read_head, out, ok = 0, 0, [], True
        while ri < rlen - K + 1:
            wl, usable = R.window_len(ri)
            if not usable:
                ri += wl
                continue
            k = key_for(ri, wl, read_head)
            if k not in cache:
                ok = False
                collect.add((ri, wl, read_head, k))
                # optimistic continuation: every event consumed, last state matched
                rh, obs, ev, lo, hi = R.gather(ri, wl, read_head)
                if obs.size < 2:
                    read_head = rh
                    ri += wl
                    continue
                read_head = rh + len(ev)
                ri += wl - K + 1
                continue
            res = cache[k]
            if res[0] is None:
                read_head = max(read_head, res[1]) if k[1] is None else res[1]
                ri += wl
                continue
            recs, last_m_ref, last_m_ev, rh = res
            if ok:
                out.extend(recs)
            read_head = rh + last_m_ev + 1
            ri += last_m_ref + 1
        return ok, out

    while True:
        need = set()
        ok, out = walk(need)
        if ok:
            return out, rounds
        rounds.append(len(need))
        for ri, wl, read_head, k in need:          # <- every one of these is independent: the parallel launch
            cache[k] = run_window(R, ri, wl, read_head)


def check(R):
    """run the speculative algorithm on R and compare with the serial chain's records; returns windows run per round"""
    out, rounds = speculative(R)
    got = np.asarray(out, dtype=np.int64).reshape(-1, 4)
    s = R.serial
    assert np.array_equal(got[:, 0], s["event"]) and np.array_equal(got[:, 1], s["ref_pos"])
    assert np.array_equal(got[:, 2], s["label"]) and np.array_equal(got[:, 3], s["indel"])
    return rounds


def main():
    mean = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden",
                                "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)
    P = portbind.Port()
    ref = synth.make_reference(400_000, 5)
    reads = synth.simulate_batch(ref, [9000, 15000, 6000, 12000], mean, seed=6, sub_rate=0.01)
    for n, sr in enumerate(reads):
        R = Read.from_synth(P, sr, mean)
        rounds = check(R)
        print(f"read {n}: {len(sr.refseq)} bases, {R.serial['event'].size} records identical to the serial chain; "
              f"windows per round {rounds}")


if __name__ == "__main__":
    main()
