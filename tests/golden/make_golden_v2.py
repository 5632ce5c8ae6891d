"""Second fixture set, again produced by the UNMODIFIED reference (oracle/_ref); run in the build container only:

    python tests/golden/make_golden_v2.py      ->  tests/golden/reads_v2.npz

  a0, a1   BASELINE.json configs[3]: BrdU-substituted reads (a fraction of the T's of the read emit from
           r10.4.1_BrdU_gaussian.model), forward and reverse strand: every normaliseEvents output plus the
           llAcrossRead log-likelihood ratios (src/detect.cpp:393-574) and the rows of the unlabelled / BrdU
           tables the read touches
  i0, i1   reads whose CIGAR has insertions, deletions and soft clips, so that queryToRef (parseCigar,
           src/htsInterface.cpp:59-157) has gaps and out-of-range entries: inputs as the reference's read
           constructor produced them, and every normaliseEvents output
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from dnascent_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def store(d, p, sr, rr, o, R):
    st, ln, mn, sd = R.detect_events(sr.raw)
    d[p + "seq_bam"] = np.frombuffer(sr.seq_bam, dtype=np.uint8)
    d[p + "flag"] = np.array(sr.flag)
    d[p + "pos"] = np.array(sr.pos)
    d[p + "cigar"] = sr.cigar
    d[p + "dac"] = sr.dac
    d[p + "basecall"] = np.frombuffer(rr.basecall, dtype=np.uint8)
    d[p + "refseq"] = np.frombuffer(rr.refseq, dtype=np.uint8)
    d[p + "query_to_ref"] = rr.query_to_ref
    d[p + "et_n"] = np.array(mn.size)
    d[p + "event_mean"] = o["event_mean"].astype(np.float32)
    assert np.all(d[p + "event_mean"].astype(np.float64) == o["event_mean"])
    d[p + "event_raw_len"] = o["event_raw_len"]
    d[p + "align"] = np.stack([o["align_event"], o["align_kmer"]], axis=1).astype(np.uint32)
    d[p + "scalars"] = np.array([o["shift"], o["scale"], o["events_per_base"], o["rough_shift"], o["rough_scale"],
                                 o["avg_log_emission"], float(o["spanned"]), float(o["max_gap"])])
    d[p + "cleaned_signal"] = o["cleaned_signal"]
    d[p + "cleaned_rank"] = o["cleaned_rank"]


def indel_read(ref, start, ref_len, reverse, pore_mean, rng, name):
    """Query = reference slice with random insertions / deletions and a soft clip at both ends."""
    rs = np.frombuffer(ref[start:start + ref_len], dtype=np.uint8)
    q, ops = [], []

    def push(op, n):
        if n == 0:
            return
        if ops and ops[-1][0] == op:
            ops[-1][1] += n
        else:
            ops.append([op, n])

    clip5, clip3 = int(rng.integers(3, 12)), int(rng.integers(3, 12))
    q.append(synth.BASES[rng.integers(0, 4, size=clip5)])
    push(4, clip5)
    i = 0
    while i < ref_len:
        run = int(min(rng.integers(40, 400), ref_len - i))
        q.append(rs[i:i + run])
        push(0, run)
        i += run
        if i >= ref_len:
            break
        if rng.random() < 0.5:
            n = int(rng.integers(1, 5))
            q.append(synth.BASES[rng.integers(0, 4, size=n)])
            push(1, n)
        else:
            n = int(min(rng.integers(1, 6), ref_len - i - 1))
            push(2, n)
            i += n
    q.append(synth.BASES[rng.integers(0, 4, size=clip3)])
    push(4, clip3)
    seq_bam = np.concatenate(q).tobytes()
    cigar = np.array([(n << 4) | op for op, n in ops], dtype=np.uint32)
    basecall = synth.revcomp(seq_bam) if reverse else seq_bam
    dac = synth.simulate_signal(basecall, pore_mean, rng)
    return synth.SynthRead(name=name, seq_bam=seq_bam, flag=16 if reverse else 0, pos=start, cigar=cigar, basecall=basecall,
                           refseq=b"", query_to_ref=np.zeros(0, dtype=np.int32), dac=dac, raw=synth.dac_to_pa(dac))


def main():
    R = refbind.Ref()
    R.configure_from_files()
    pore_mean, _ = R.get_model(refbind.PORE)
    unl_m, unl_s = R.get_model(refbind.UNLABELLED)
    ana_m, ana_s = R.get_model(refbind.ANALOGUE)
    ref = synth.make_reference(40_000, seed=20240701)
    R.set_reference(ref)
    rng = np.random.default_rng(20240702)
    d = {"reference": np.frombuffer(ref, dtype=np.uint8)}

    # ---- analogue-substituted reads (configs[3]) ----
    touched = []
    for i, (L, rev, frac) in enumerate([(3500, False, 0.5), (3200, True, 0.8)]):
        start = int(rng.integers(0, len(ref) - L))
        sl = ref[start:start + L]
        basecall = synth.revcomp(sl) if rev else sl
        ranks = synth.kmer_ranks(basecall)
        seq = np.frombuffer(basecall, dtype=np.uint8)
        sub = (seq == ord("T")) & (rng.random(seq.size) < frac)
        mask = sub[4:4 + ranks.size] & (ana_s[ranks] > 0)          # centre base is a substituted T
        sr = synth.simulate_read(ref, start, L, rev, pore_mean, rng, name=f"a{i}",
                                 level_override=(mask, ana_m[ranks], ana_s[ranks]))
        rr = R.read_new(sr)
        o = rr.normalise(staged=True)
        store(d, f"a{i}_", sr, rr, o, R)
        pos, llr = rr.ll_across_read(12)
        d[f"a{i}_pos_global"], d[f"a{i}_llr"] = pos, llr
        d[f"a{i}_brdu_fraction"] = np.array(frac)
        touched.append(np.unique(synth.kmer_ranks(rr.refseq)))
        print(f"a{i}: L={L} rev={rev} f={frac} align={o['align_event'].size} calls={pos.size} "
              f"mean LLR={llr.mean():.3f} (substituted k-mers {int(mask.sum())})")
    ranks = np.unique(np.concatenate(touched)).astype(np.uint32)
    d.update(ranks=ranks, unl_mean=unl_m[ranks], unl_stdv=unl_s[ranks], ana_mean=ana_m[ranks], ana_stdv=ana_s[ranks])

    # ---- indel reads ----
    for i, (L, rev) in enumerate([(3000, False), (2600, True)]):
        sr = indel_read(ref, int(rng.integers(0, len(ref) - L - 10)), L, rev, pore_mean, rng, f"i{i}")
        rr = R.read_new(sr)
        o = rr.normalise(staged=True)
        store(d, f"i{i}_", sr, rr, o, R)
        q2r = rr.query_to_ref
        print(f"i{i}: ref_len={L} rev={rev} query={len(rr.basecall)} cigar_ops={sr.cigar.size} "
              f"q2r gaps={(q2r < 0).sum()} align={o['align_event'].size} cleaned={o['cleaned_signal'].size}")
    np.savez_compressed(os.path.join(OUT, "reads_v2.npz"), **d)
    print("reads_v2.npz", os.path.getsize(os.path.join(OUT, "reads_v2.npz")))


if __name__ == "__main__":
    main()
