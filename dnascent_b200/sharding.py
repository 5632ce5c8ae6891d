"""Length-balanced partitioning of reads: across GPUs (ranks) and, within a GPU, into device-sized bins.

Reads are independent units (the reference parallelises the same way: `#pragma omp parallel for schedule(dynamic)`
over reads, /root/reference/src/detect.cpp:852), and one read can never be split -- the band placement of the
alignment depends on every previous band.  So multi-GPU is pure sharding: no collective on the data path, results
are gathered on the host.  Cost model: device time is proportional to the number of samples (events and k-mers,
hence DP cells, scale with it), so bins are balanced on n_samples.
"""
from __future__ import annotations

import heapq
import numpy as np


def shard_reads(n_samples, n_ranks: int) -> list[np.ndarray]:
    """Longest-processing-time-first assignment of read indices to ranks; deterministic.

    Returns, per rank, the (ascending) indices of the reads it owns.  max load <= (4/3 - 1/(3m)) * optimum, and in
    practice within one read of perfectly even for read-length distributions.
    """
    n_samples = np.asarray(n_samples, dtype=np.int64)
    order = np.argsort(-n_samples, kind="stable")
    heap = [(0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    owner = np.empty(n_samples.size, dtype=np.int32)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + int(n_samples[i]), r))
    return [np.flatnonzero(owner == r) for r in range(n_ranks)]


def make_bins(n_samples, max_samples_per_bin: int, max_reads_per_bin: int = 1 << 20, policy: str = "sorted") -> list[np.ndarray]:
    """Bins of whole reads for one GPU (one bin = one device launch sequence).  Each bin lists its reads longest first.

    A read's band chain is serial (one warp walks it), so a launch lasts at least as long as its longest read's chain
    (~0.4 s for a 2.7*10^6-sample read on a busy SM, about what a whole 2.5*10^9-sample bin takes at full throughput).
    policy "sorted" (default): consecutive runs of the length-sorted list -- every bin holds reads of similar length,
    only the first bin is bound by the long chains, the others run at full throughput.
    policy "interleaved": reads dealt round-robin to ceil(total / budget) bins so that every bin has the same length
    mix.  Measured on a B200 (30 000 reads of the C2 law, 4 bins, profiles/r2m_interleaved.json): 4 897 Msamples/s
    against 7 059 for "sorted" -- every launch then waits for a long chain.  Kept for workloads whose bins are large
    against their longest read."""
    n_samples = np.asarray(n_samples, dtype=np.int64)
    order = np.argsort(-n_samples, kind="stable")
    if policy == "sorted":
        bins, cur, load = [], [], 0
        for i in order:
            s = int(n_samples[i])
            if cur and (load + s > max_samples_per_bin or len(cur) >= max_reads_per_bin):
                bins.append(np.array(cur, dtype=np.int64))
                cur, load = [], 0
            cur.append(int(i))
            load += s
        if cur:
            bins.append(np.array(cur, dtype=np.int64))
        return bins
    if policy != "interleaved":
        raise ValueError(policy)
    if order.size == 0:
        return []
    total = int(n_samples.sum())
    nb = max(1, -(-total // max(int(max_samples_per_bin), 1)), -(-order.size // max_reads_per_bin))
    while True:
        bins = [order[j::nb] for j in range(nb)]
        # round-robin over a descending list: bin 0 is the heaviest; grow the bin count until it fits the budget
        # (a single read larger than the budget gets a bin of its own size: it cannot be split)
        heaviest = int(n_samples[bins[0]].sum())
        if heaviest <= max_samples_per_bin or nb >= order.size or heaviest == int(n_samples[order[0]]):
            return [b for b in bins if b.size]
        nb += 1


def bin_budget_samples(free_bytes: int, resident_bytes_per_sample: float = 2.6,
                       workspace_bytes_per_sample: float = 27.0, n_resident_bins: int = 1) -> int:
    """How many samples fit a bin: the inputs of every resident bin stay in HBM (int16 DAC + sequences ~2.6 B/sample),
    one bin's workspace is live at a time (event slots, trace rows, alignment slots ~27 B/sample)."""
    usable = int(free_bytes * 0.9)
    return max(int(usable / (resident_bytes_per_sample * n_resident_bins + workspace_bytes_per_sample)), 1 << 20)
