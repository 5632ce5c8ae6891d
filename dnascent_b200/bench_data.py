"""Device-side generation of PromethION-scale synthetic workloads for bench.py.

Same recipe as synth.py (random reference, reads are exact substrings on a random strand, per 9-mer dwell
3 + Geom(1/10.5), level shift + scale * (mu + 0.14 N(0,1)), int16 DAC quantisation) but vectorised over millions
of reads with torch on the GPU, because 3*10^10 samples cannot be produced with a Python loop.  torch is used here
as a random-number and gather engine for INPUT DATA only; nothing in this module is on the measured path.
"""
from __future__ import annotations

import dataclasses
import numpy as np

from . import _lib
from .synth import DAC_OFFSET, DAC_SCALE, lognormal_lengths, make_reference  # noqa: F401

K = 9


@dataclasses.dataclass
class Workload:
    """Host-resident reads in concatenated form (what a POD5/BAM loader would hand over)."""
    dac: np.ndarray          # int16, all reads back to back
    raw_off: np.ndarray      # int64 [R+1]
    seq: np.ndarray          # uint8 ASCII, basecall == reference slice (exact-match reads), sequencing orientation
    seq_off: np.ndarray      # int64 [R+1]
    q2r: np.ndarray          # int32 arange(max read length): the dense queryToRef of every `{L}M` read
    runs: np.ndarray = None  # dnb_q2r_run [R]: the same map as one run per read (what parseCigar gives for `{L}M`)

    @property
    def n_reads(self) -> int:
        return self.raw_off.size - 1

    @property
    def n_samples(self) -> np.ndarray:
        return np.diff(self.raw_off)

    def descs(self, idx=None, dense_q2r: bool = False) -> np.ndarray:
        """dnb_read_desc array (numpy mirror) for the reads `idx`, pointing into this workload's buffers.  queryToRef
        goes as one run per read (16 B) unless dense_q2r asks for the int32-per-base form."""
        if idx is None:
            idx = np.arange(self.n_reads)
        idx = np.asarray(idx, dtype=np.int64)
        d = np.zeros(idx.size, dtype=_lib.READ_DESC_DTYPE)
        d["raw_pA"] = 0
        d["raw_dac"] = self.dac.ctypes.data + 2 * self.raw_off[idx]
        d["dac_offset"] = float(DAC_OFFSET)
        d["dac_scale"] = float(DAC_SCALE)
        d["n_samples"] = self.raw_off[idx + 1] - self.raw_off[idx]
        sl = (self.seq_off[idx + 1] - self.seq_off[idx]).astype(np.uint32)
        d["query"] = self.seq.ctypes.data + self.seq_off[idx]
        d["query_len"] = sl
        d["ref"] = d["query"]
        d["ref_len"] = sl
        if dense_q2r:
            d["query_to_ref"] = self.q2r.ctypes.data
        else:
            if self.runs is None:
                self.runs = np.zeros(self.n_reads, dtype=_lib.Q2R_RUN_DTYPE)
                self.runs["len"] = np.diff(self.seq_off)
                self.runs["stride"] = 1
            d["q2r_runs"] = self.runs.ctypes.data + self.runs.itemsize * idx
            d["n_q2r_runs"] = 1
        return d

    def read(self, i: int):
        """One read as plain arrays (for the CPU baseline / parity checks)."""
        from .synth import dac_to_pa
        dac = self.dac[self.raw_off[i]:self.raw_off[i + 1]]
        seq = self.seq[self.seq_off[i]:self.seq_off[i + 1]].tobytes()
        return dict(dac=dac, raw=dac_to_pa(dac), basecall=seq, refseq=seq, query_to_ref=self.q2r[:len(seq)])


def generate(lengths, model_mean: np.ndarray, seed: int, device: str = "cuda:0", ref_len: int = 1_000_000,
             dwell_p: float = 1.0 / 10.5, chunk_bases: int = 40_000_000) -> Workload:
    import torch

    dev = torch.device(device)
    lengths = np.minimum(np.asarray(lengths, dtype=np.int64), ref_len - 1)
    R = lengths.size
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ref_codes = torch.randint(0, 4, (ref_len,), generator=g, device=dev, dtype=torch.int64)   # 0..3 = A C G T
    ascii_of = torch.tensor([65, 67, 71, 84], device=dev, dtype=torch.uint8)
    rank_of = torch.tensor([0, 3, 2, 1], device=dev, dtype=torch.int64)    # reference alphabet A=0 T=1 G=2 C=3
    mean_t = torch.as_tensor(model_mean, device=dev, dtype=torch.float32)
    pw = (4 ** torch.arange(K - 1, -1, -1, device=dev, dtype=torch.int64))

    dac_parts, seq_parts, ns_parts = [], [], []
    r0 = 0
    while r0 < R:
        r1 = r0
        tot = 0
        while r1 < R and (tot == 0 or tot + lengths[r1] <= chunk_bases):
            tot += int(lengths[r1])
            r1 += 1
        L = torch.as_tensor(lengths[r0:r1], device=dev)
        n = r1 - r0
        boff = torch.zeros(n + 1, device=dev, dtype=torch.int64)
        boff[1:] = torch.cumsum(L, 0)
        start = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (ref_len - L).double()).long()
        rev = torch.rand(n, generator=g, device=dev) < 0.5
        rid = torch.repeat_interleave(torch.arange(n, device=dev), L)
        pos = torch.arange(tot, device=dev) - boff[rid]
        refpos = torch.where(rev[rid], start[rid] + L[rid] - 1 - pos, start[rid] + pos)
        code = ref_codes[refpos]
        code = torch.where(rev[rid], 3 - code, code)                       # complement in A C G T order
        seq_parts.append(ascii_of[code].cpu().numpy())
        rc = torch.cat([rank_of[code], torch.zeros(K, device=dev, dtype=torch.int64)])
        rank = torch.zeros(tot, device=dev, dtype=torch.int64)
        for j in range(K):
            rank += rc[j:j + tot] * pw[j]
        valid = pos <= (L[rid] - K)
        krid = rid[valid]
        mu = mean_t[rank[valid]]
        del rank, rc, code, refpos, pos
        nk = mu.numel()
        dwell = 2 + torch.empty(nk, device=dev, dtype=torch.float32).geometric_(dwell_p, generator=g).long()
        shift = 90.0 + 5.0 * torch.randn(n, generator=g, device=dev)
        scale = 15.0 + 1.5 * torch.randn(n, generator=g, device=dev)
        a = shift[krid] + scale[krid] * mu
        bq = scale[krid] * 0.14
        ns = torch.zeros(n, device=dev, dtype=torch.int64).index_add_(0, krid, dwell)
        a_s = torch.repeat_interleave(a, dwell)
        b_s = torch.repeat_interleave(bq, dwell)
        pa = a_s + b_s * torch.randn(a_s.numel(), generator=g, device=dev)
        dac = torch.round(pa / float(DAC_SCALE) - float(DAC_OFFSET)).clamp_(-32768, 32767).to(torch.int16)
        dac_parts.append(dac.cpu().numpy())
        ns_parts.append(ns.cpu().numpy())
        del a_s, b_s, pa, dac, a, bq, dwell, mu, krid, rid, valid
        r0 = r1
    n_samples = np.concatenate(ns_parts)
    raw_off = np.zeros(R + 1, dtype=np.int64)
    raw_off[1:] = np.cumsum(n_samples)
    seq_off = np.zeros(R + 1, dtype=np.int64)
    seq_off[1:] = np.cumsum(lengths)
    torch.cuda.empty_cache()

    def join(parts, total, dtype):
        # assemble without ever holding two full copies (np.empty pages are not resident until written)
        out = np.empty(total, dtype=dtype)
        pos = 0
        for i in range(len(parts)):
            p = parts[i]
            out[pos:pos + p.size] = p
            pos += p.size
            parts[i] = None
        return out

    return Workload(dac=join(dac_parts, int(raw_off[-1]), np.int16), raw_off=raw_off,
                    seq=join(seq_parts, int(seq_off[-1]), np.uint8), seq_off=seq_off,
                    q2r=np.arange(int(lengths.max()) if R else 1, dtype=np.int32))
