"""Third fixture set, produced by the UNMODIFIED reference (oracle/_ref); run in the build container only:

    python tests/golden/make_golden_v3.py      ->  tests/golden/reads_v3.npz

  e0, e1   BASELINE.json configs[3], EdU: reads in which a fraction of the T's emit from r10.4.1_EdU_gaussian.model,
           scored by the reference's llAcrossRead / sequenceProbability with Pore_Substrate_Config.analogue_model
           re-pointed at that file through the reference's own parser (SURVEY s.0.2: src/config.h:50,54 only ever load
           the BrdU table; the EdU one ships in pore_models/ and has the same format), forward and reverse strand
  j0, j1   reads with insertions / deletions / soft clips scored with the BrdU table: refToQuery then has runs of equal
           values, which is what llAcrossRead's readHead scan is sensitive to (src/detect.cpp:446-512)
Per read: the reference's normaliseEvents outputs, refToQuery, and the (position, LLR) calls of llAcrossRead(r, 12);
plus the rows of the unlabelled / BrdU / EdU tables these reads touch (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import refbind  # noqa: E402
from dnascent_b200 import synth  # noqa: E402
from make_golden_v2 import store, indel_read  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    R = refbind.Ref()
    R.configure_from_files()
    pore_mean, _ = R.get_model(refbind.PORE)
    unl_m, unl_s = R.get_model(refbind.UNLABELLED)
    brdu_m, brdu_s = R.get_model(refbind.ANALOGUE)
    ref = synth.make_reference(40_000, seed=20251001)
    R.set_reference(ref)
    rng = np.random.default_rng(20251002)
    d = {"reference": np.frombuffer(ref, dtype=np.uint8)}
    touched = []

    # ---- indel reads, BrdU table ----
    for i, (L, rev) in enumerate([(3000, False), (2800, True)]):
        sr = indel_read(ref, int(rng.integers(0, len(ref) - L - 10)), L, rev, pore_mean, rng, f"j{i}")
        rr = R.read_new(sr)
        o = rr.normalise(staged=True)
        store(d, f"j{i}_", sr, rr, o, R)
        d[f"j{i}_ref_to_query"] = rr.ref_to_query
        pos, llr = rr.ll_across_read(12)
        d[f"j{i}_pos_global"], d[f"j{i}_llr"] = pos, llr
        d[f"j{i}_ref_start"], d[f"j{i}_ref_end"] = np.array(rr.ref_start), np.array(rr.ref_end)
        touched.append(np.unique(synth.kmer_ranks(rr.refseq)))
        print(f"j{i}: ref_len={L} rev={rev} query={len(rr.basecall)} align={o['align_event'].size} calls={pos.size}")

    # ---- EdU: analogue_model <- r10.4.1_EdU_gaussian.model, parsed by the reference ----
    R.load_model_file(refbind.ANALOGUE, "r10.4.1_EdU_gaussian.model", fit_stdv=True)
    edu_m, edu_s = R.get_model(refbind.ANALOGUE)
    assert not np.array_equal(edu_m, brdu_m)
    for i, (L, rev, frac) in enumerate([(3400, False, 0.5), (3100, True, 0.8)]):
        start = int(rng.integers(0, len(ref) - L))
        sl = ref[start:start + L]
        basecall = synth.revcomp(sl) if rev else sl
        ranks = synth.kmer_ranks(basecall)
        seq = np.frombuffer(basecall, dtype=np.uint8)
        sub = (seq == ord("T")) & (rng.random(seq.size) < frac)
        mask = sub[4:4 + ranks.size] & (edu_s[ranks] > 0)          # centre base is a substituted T
        sr = synth.simulate_read(ref, start, L, rev, pore_mean, rng, name=f"e{i}",
                                 level_override=(mask, edu_m[ranks], edu_s[ranks]))
        rr = R.read_new(sr)
        o = rr.normalise(staged=True)
        store(d, f"e{i}_", sr, rr, o, R)
        d[f"e{i}_ref_to_query"] = rr.ref_to_query
        pos, llr = rr.ll_across_read(12)
        d[f"e{i}_pos_global"], d[f"e{i}_llr"] = pos, llr
        d[f"e{i}_ref_start"], d[f"e{i}_ref_end"] = np.array(rr.ref_start), np.array(rr.ref_end)
        d[f"e{i}_edu_fraction"] = np.array(frac)
        touched.append(np.unique(synth.kmer_ranks(rr.refseq)))
        print(f"e{i}: L={L} rev={rev} f={frac} align={o['align_event'].size} calls={pos.size} mean LLR={llr.mean():.3f} "
              f"(substituted k-mers {int(mask.sum())})")
    ranks = np.unique(np.concatenate(touched)).astype(np.uint32)
    d.update(ranks=ranks, unl_mean=unl_m[ranks], unl_stdv=unl_s[ranks], brdu_mean=brdu_m[ranks], brdu_stdv=brdu_s[ranks],
             edu_mean=edu_m[ranks], edu_stdv=edu_s[ranks])
    np.savez_compressed(os.path.join(OUT, "reads_v3.npz"), **d)
    print("reads_v3.npz", os.path.getsize(os.path.join(OUT, "reads_v3.npz")))


if __name__ == "__main__":
    main()
