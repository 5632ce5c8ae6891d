"""Seeded synthetic R10.4.1 reads for parity tests and benchmarks.

Nothing here is on the hot path: it only manufactures inputs.  A read is an exact substring of a
random reference (optionally reverse strand, optionally with substitutions so that query != reference),
its signal is drawn per 9-mer from the ONT r10.4.1_400bps level table: dwell ``3 + Geom(p)`` samples,
level ``shift + scale * (mu_kmer + 0.14 * N(0,1))`` pA, quantised to an int16 DAC value and converted back
with the exact float32 expression shape of the reference's POD5 reader
(``((float)dac + (float)offset) * (float)scale``, /root/reference/src/pod5.cpp:60).
"""
from __future__ import annotations

import dataclasses
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
# reference alphabet order A=0,T=1,G=2,C=3 (/root/reference/src/data_IO.cpp:131)
_RANK_OF = np.zeros(256, dtype=np.int64)
_RANK_OF[ord("A")] = 0
_RANK_OF[ord("T")] = 1
_RANK_OF[ord("G")] = 2
_RANK_OF[ord("C")] = 3
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b

DAC_OFFSET = np.float32(-240.0)
DAC_SCALE = np.float32(0.1465)
K = 9


def make_reference(n_bases: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    return BASES[rng.integers(0, 4, size=n_bases)].tobytes()


def revcomp(seq: bytes) -> bytes:
    return _COMP[np.frombuffer(seq, dtype=np.uint8)][::-1].tobytes()


def kmer_ranks(seq: bytes, k: int = K) -> np.ndarray:
    """Base-4 rank of every k-mer, first base most significant (kmer2index, data_IO.cpp:129-141)."""
    d = _RANK_OF[np.frombuffer(seq, dtype=np.uint8)]
    n = len(seq) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint32)
    r = np.zeros(n, dtype=np.int64)
    for j in range(k):
        r = r * 4 + d[j:j + n]
    return r.astype(np.uint32)


@dataclasses.dataclass
class SynthRead:
    name: str
    seq_bam: bytes          # SEQ column (reference-strand orientation)
    flag: int               # 0 or 16
    pos: int                # 0-based leftmost reference coordinate
    cigar: np.ndarray       # BAM-encoded uint32 ops
    basecall: bytes         # sequencing orientation (what normaliseEvents sees)
    refseq: bytes           # reference slice in sequencing orientation
    query_to_ref: np.ndarray  # dense int32, -1 = absent
    dac: np.ndarray         # int16
    raw: np.ndarray         # float32 pA


def dac_to_pa(dac: np.ndarray) -> np.ndarray:
    return (dac.astype(np.float32) + DAC_OFFSET) * DAC_SCALE


def simulate_signal(seq: bytes, model_mean: np.ndarray, rng: np.random.Generator, dwell_p: float = 1.0 / 10.5,
                    shift: float | None = None, scale: float | None = None, noise_sd: float = 0.14,
                    level_override: tuple[np.ndarray, np.ndarray, np.ndarray] | None = None) -> np.ndarray:
    """int16 DAC samples for `seq` (sequencing orientation).

    level_override = (mask over k-mers, mean, stdv): k-mers whose mask is set draw from that Gaussian
    instead of the ONT table (analogue-substituted reads, App. D item 4).
    """
    ranks = kmer_ranks(seq)
    mu = model_mean[ranks].astype(np.float64)
    sd = np.full(mu.shape, noise_sd)
    if level_override is not None:
        mask, omu, osd = level_override
        mu = np.where(mask, omu, mu)
        sd = np.where(mask, osd, sd)
    if shift is None:
        shift = rng.normal(90.0, 5.0)
    if scale is None:
        scale = rng.normal(15.0, 1.5)
    dwell = 3 + (rng.geometric(dwell_p, size=mu.size) - 1)
    lvl = np.repeat(mu, dwell)
    sdr = np.repeat(sd, dwell)
    pa = shift + scale * (lvl + sdr * rng.standard_normal(lvl.size))
    dac = np.rint(pa / float(DAC_SCALE) - float(DAC_OFFSET))
    return np.clip(dac, -32768, 32767).astype(np.int16)


def simulate_read(ref: bytes, start: int, length: int, reverse: bool, model_mean: np.ndarray,
                  rng: np.random.Generator, name: str = "read", sub_rate: float = 0.0,
                  dwell_p: float = 1.0 / 10.5, **sig_kw) -> SynthRead:
    """One read: CIGAR `{L}M`; with sub_rate>0 the basecall carries substitutions relative to the reference."""
    ref_slice = ref[start:start + length]
    q = np.frombuffer(ref_slice, dtype=np.uint8).copy()
    if sub_rate > 0:
        m = rng.random(q.size) < sub_rate
        q[m] = BASES[(np.searchsorted(BASES, q[m]) + rng.integers(1, 4, size=int(m.sum()))) % 4]
    seq_bam = q.tobytes()
    if reverse:
        basecall, refseq = revcomp(seq_bam), revcomp(ref_slice)
    else:
        basecall, refseq = seq_bam, ref_slice
    dac = simulate_signal(basecall, model_mean, rng, dwell_p=dwell_p, **sig_kw)
    cigar = np.array([(length << 4) | 0], dtype=np.uint32)
    q2r = np.arange(length, dtype=np.int32)  # parseCigar for `{L}M` (htsInterface.cpp:59-157), either strand
    return SynthRead(name=name, seq_bam=seq_bam, flag=16 if reverse else 0, pos=start, cigar=cigar,
                     basecall=basecall, refseq=refseq, query_to_ref=q2r, dac=dac, raw=dac_to_pa(dac))


def lognormal_lengths(n: int, n50: float, rng: np.random.Generator, sigma: float = 0.6,
                      lo: int = 1000, hi: int = 1_000_000) -> np.ndarray:
    """Read lengths whose N50 is `n50`: for a log-normal, N50 = exp(mu + sigma^2) (length-weighted median)."""
    mu = np.log(n50) - sigma * sigma
    return np.clip(rng.lognormal(mu, sigma, size=n), lo, hi).astype(np.int64)


def simulate_batch(ref: bytes, lengths, model_mean: np.ndarray, seed: int, sub_rate: float = 0.0,
                   dwell_p: float = 1.0 / 10.5) -> list[SynthRead]:
    rng = np.random.default_rng(seed)
    out = []
    for i, L in enumerate(lengths):
        L = int(min(L, len(ref) - 1))
        start = int(rng.integers(0, len(ref) - L))
        out.append(simulate_read(ref, start, L, bool(rng.integers(0, 2)), model_mean, rng, name=f"read{i}",
                                 sub_rate=sub_rate, dwell_p=dwell_p))
    return out
