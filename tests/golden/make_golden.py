"""Regenerates the committed fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Fixtures (all small, all produced by the reference's own code):
  pore_model_r10.4.1_400bps.npz   pore_model means exactly as import_poreModel_staticStdv loads them
                                  (src/data_IO.cpp:144-186); they are float32-exact, stored as float32
  reads_v1.npz                    6 seeded synthetic reads (inputs included) with every output of the reference's
                                  normaliseEvents (src/event_handling.cpp:544-607) and detect_events
  hmm_v1.npz                      sequenceProbability / llAcrossRead goldens (src/detect.cpp:235-574) for one read,
                                  with the (mean, stdv) rows of the unlabelled and BrdU tables that the sites touch
  probability_v1.npz              eexp/eln/lnSum/lnProd/lnGreaterThan/*PDF known answers (src/probability.cpp)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from dnascent_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    R = refbind.Ref()
    R.configure_from_files()
    pore_mean, pore_sd = R.get_model(refbind.PORE)
    assert np.all(pore_sd == 0.14)
    assert np.all(pore_mean.astype(np.float32).astype(np.float64) == pore_mean)
    np.savez_compressed(os.path.join(OUT, "pore_model_r10.4.1_400bps.npz"), mean=pore_mean.astype(np.float32))

    # ---- reads ----
    ref = synth.make_reference(60_000, seed=20240601)
    R.set_reference(ref)
    rng = np.random.default_rng(20240602)
    spec = [(1300, False, 0.0), (2100, True, 0.0), (3000, False, 0.02), (1800, True, 0.03), (4000, False, 0.0),
            (1050, True, 0.0)]
    d = {"reference": np.frombuffer(ref, dtype=np.uint8), "n_reads": np.array(len(spec))}
    for i, (L, rev, sub) in enumerate(spec):
        sr = synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, rev, pore_mean, rng, name=f"g{i}", sub_rate=sub)
        rr = R.read_new(sr)
        o = rr.normalise(staged=True)
        st, ln, mn, sd = R.detect_events(sr.raw)
        p = f"r{i}_"
        d[p + "seq_bam"] = np.frombuffer(sr.seq_bam, dtype=np.uint8)
        d[p + "flag"] = np.array(sr.flag)
        d[p + "pos"] = np.array(sr.pos)
        d[p + "cigar"] = sr.cigar
        d[p + "dac"] = sr.dac
        d[p + "basecall"] = np.frombuffer(rr.basecall, dtype=np.uint8)
        d[p + "refseq"] = np.frombuffer(rr.refseq, dtype=np.uint8)
        d[p + "query_to_ref"] = rr.query_to_ref
        d[p + "et_start"] = st.astype(np.uint32)
        d[p + "et_length"] = ln
        d[p + "et_mean"] = mn
        d[p + "et_stdv"] = sd
        d[p + "event_mean"] = o["event_mean"].astype(np.float32)
        assert np.all(d[p + "event_mean"].astype(np.float64) == o["event_mean"])
        d[p + "event_raw_len"] = o["event_raw_len"]
        d[p + "align"] = np.stack([o["align_event"], o["align_kmer"]], axis=1).astype(np.uint32)
        d[p + "scalars"] = np.array([o["shift"], o["scale"], o["events_per_base"], o["rough_shift"], o["rough_scale"],
                                     o["avg_log_emission"], float(o["spanned"]), float(o["max_gap"])])
        d[p + "cleaned_signal"] = o["cleaned_signal"]
        d[p + "cleaned_rank"] = o["cleaned_rank"]
        print(f"read {i}: L={L} rev={rev} samples={sr.raw.size} events={o['event_mean'].size} "
              f"align={o['align_event'].size} shift={o['shift']:.4f} scale={o['scale']:.4f}")
        if i == 4:
            hmm_read, hmm_out, hmm_sr = rr, o, sr
    np.savez_compressed(os.path.join(OUT, "reads_v1.npz"), **d)

    # ---- HMM ----
    unl_m, unl_s = R.get_model(refbind.UNLABELLED)
    ana_m, ana_s = R.get_model(refbind.ANALOGUE)
    pos, llr = hmm_read.ll_across_read(12)
    refseq = hmm_read.refseq
    r2q = hmm_read.ref_to_query
    # every k-mer rank the sites can touch
    ranks = np.unique(synth.kmer_ranks(refseq))
    h = dict(read_index=np.array(4), pos_global=pos, llr=llr, ref_start=np.array(R.L.dnbref_read_ref_start(hmm_read.h)),
             ref_end=np.array(R.L.dnbref_read_ref_end(hmm_read.h)), is_reverse=np.array(R.L.dnbref_read_is_reverse(hmm_read.h)),
             ref_to_query=r2q, ranks=ranks.astype(np.uint32), unl_mean=unl_m[ranks], unl_stdv=unl_s[ranks],
             ana_mean=ana_m[ranks], ana_stdv=ana_s[ranks])
    # direct sequenceProbability known answers on hand-made snippets
    rng2 = np.random.default_rng(5)
    sp_obs, sp_seq, sp_par, sp_out = [], [], [], []
    for j in range(40):
        start = int(rng2.integers(0, len(refseq) - 40))
        sn = refseq[start:start + 33]
        n_obs = int(rng2.integers(15, 70))
        kr = synth.kmer_ranks(sn)
        lv = np.repeat(unl_m[kr[:24]], 3)[:n_obs] if n_obs <= 72 else None
        shift, scale = float(rng2.normal(90, 5)), float(rng2.normal(15, 1.5))
        obs = shift + scale * (lv + 0.12 * rng2.standard_normal(lv.size))
        epb = float(rng2.uniform(1.6, 2.6))
        la = R.sequence_probability(obs, sn, 12, True, shift, scale, epb, 8, 16)
        lt = R.sequence_probability(obs, sn, 12, False, shift, scale, epb, 0, 0)
        sp_obs.append(obs); sp_seq.append(np.frombuffer(sn, dtype=np.uint8)); sp_par.append([shift, scale, epb]); sp_out.append([la, lt])
    h["sp_obs"] = np.concatenate(sp_obs)
    h["sp_obs_off"] = np.cumsum([0] + [len(o) for o in sp_obs]).astype(np.uint64)
    h["sp_seq"] = np.stack(sp_seq)
    h["sp_par"] = np.array(sp_par)
    h["sp_out"] = np.array(sp_out)
    np.savez_compressed(os.path.join(OUT, "hmm_v1.npz"), **h)
    print(f"hmm: {pos.size} calls on read 4, {len(sp_out)} direct cases, {ranks.size} table rows")

    # ---- probability.cpp known answers ----
    nan = float("nan")
    xs = [nan, -745.2, -30.5, -1.25, 0.0, 1e-300, 0.5, 1.0, 3.75, 700.0]
    pr = dict(xs=np.array(xs), eexp=np.array([R.L.dnbref_eexp(x) for x in xs]))
    eln = []
    for x in xs:
        try:
            eln.append(R.eln(x) if not np.isnan(x) else np.inf)   # eln(NaN) throws in the reference: marked +inf
        except ValueError:
            eln.append(np.inf)
    pr["eln"] = np.array(eln)
    pairs = [(a, b) for a in xs[:8] for b in xs[:8]]
    pr["pairs"] = np.array(pairs)
    pr["lnSum"] = np.array([R.L.dnbref_lnSum(a, b) for a, b in pairs])
    pr["lnProd"] = np.array([R.L.dnbref_lnProd(a, b) for a, b in pairs])
    pr["lnGreaterThan"] = np.array([R.L.dnbref_lnGreaterThan(a, b) for a, b in pairs])
    trip = [(0.0, 1.0, 0.3), (-1.5, 0.14, -1.2), (2.0, 0.05, 9.0), (90.0, 15.0, 70.0), (0.2, 3.0, 0.2)]
    pr["triples"] = np.array(trip)
    pr["uniformPDF"] = np.array([R.L.dnbref_uniformPDF(*t) for t in trip])
    pr["normalPDF"] = np.array([R.L.dnbref_normalPDF(*t) for t in trip])
    pr["cauchyPDF"] = np.array([R.L.dnbref_cauchyPDF(*t) for t in trip])
    np.savez_compressed(os.path.join(OUT, "probability_v1.npz"), **pr)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
