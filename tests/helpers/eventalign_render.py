"""Test helper: rebuild eventalign's humanReadable_eventalignOut (src/alignment.cpp:553, 676-736) from the record form
(event, ref_pos, label, indelScore) that the oracle port and the CUDA path produce, so that records can be compared
with the reference's text byte for byte.  std::to_string(double) is printf("%f")."""
import numpy as np

_COMP = bytes.maketrans(b"ACGT", b"TGCA")
K = 9


def render(header: bytes, refseq: bytes, ref_start: int, ref_end: int, is_reverse: bool, raw_concat: np.ndarray,
           raw_len: np.ndarray, rec: dict, shift: float, scale: float, model_mean: np.ndarray, kmer2index) -> bytes:
    off = np.concatenate([[0], np.cumsum(raw_len.astype(np.int64))])
    out = [header]
    for ev, rp, lab in zip(rec["event"].tolist(), rec["ref_pos"].tolist(), rec["label"].tolist()):
        kmer_strand = refseq[rp:rp + K]
        if is_reverse:
            coord = (ref_end - rp - K // 2 - 1) & 0xFFFFFFFF
            kmer_ref = kmer_strand.translate(_COMP)[::-1]
        else:
            coord = (ref_start + rp + K // 2) & 0xFFFFFFFF
            kmer_ref = kmer_strand
        if lab == 1:
            tail = b"\t" + kmer_strand + b"\t" + (b"%f" % model_mean[kmer2index(kmer_strand)]) + b"\n"
        else:
            tail = b"\t" + b"N" * K + b"\t0\n"
        head = b"%d\t" % coord + kmer_ref + b"\t"
        for x in raw_concat[off[ev]:off[ev + 1]].tolist():
            out.append(head + (b"%f" % ((x - shift) / scale)) + tail)
    return b"".join(out)
