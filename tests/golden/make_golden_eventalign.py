"""Fixtures for SURVEY s.8 row f1 (eventalign / builtinViterbi), produced by the UNMODIFIED reference (oracle/_ref);
run in the build container only:

    python tests/golden/make_golden_eventalign.py      ->  tests/golden/eventalign_v1.npz

  e_<tag>_*   for the reads of reads_v1.npz (g0..g5) and reads_v2.npz (a0, a1: BrdU-substituted; i0, i1: indel and
              soft-clip CIGARs): the reference's eventalign text (src/alignment.cpp:547-744) as sha256 + length + line
              count, its header line, and the same content as records (event, ref_pos, label) -- the record form is
              accepted only after re-rendering it reproduces the reference's text byte for byte
  v_*         builtinViterbi (src/alignment.cpp:193-516) known answers on 40 hand-made windows (noisy, with stalls and
              skipped k-mers so that insertion and deletion states occur): score and state path
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refbind, portbind  # noqa: E402
from dnascent_b200 import synth  # noqa: E402
from conftest import GoldenRead, GoldenReadV2  # noqa: E402
from helpers.eventalign_render import render  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    R = refbind.Ref()
    R.configure_from_files()
    P = portbind.Port()
    mean, _ = R.get_model(refbind.PORE)
    d = {}
    d1 = np.load(os.path.join(OUT, "reads_v1.npz"))
    d2 = np.load(os.path.join(OUT, "reads_v2.npz"))
    todo = [(f"g{i}", GoldenRead(d1, i), d1["reference"].tobytes()) for i in range(int(d1["n_reads"]))]
    todo += [(t, GoldenReadV2(d2, t), d2["reference"].tobytes()) for t in ("a0", "a1", "i0", "i1")]
    cur_ref = None
    for tag, g, reference in todo:
        if reference is not cur_ref:
            R.set_reference(reference)
            cur_ref = reference
        rr = R.read_new(g)
        o = rr.normalise(staged=False)
        assert np.array_equal(np.stack([o["align_event"], o["align_kmer"]], axis=1), g.align), tag
        text = rr.eventalign(50)
        rec = P.eventalign(rr.refseq, rr.ref_to_query, o["align_event"], o["align_kmer"], o["event_mean"], o["shift"],
                           o["scale"], o["events_per_base"], mean)
        header = text.split(b"\n")[0] + b"\n"
        again = render(header, rr.refseq, rr.ref_start, rr.ref_end, rr.is_reverse, rr.events_raw_concat(),
                       o["event_raw_len"], rec, o["shift"], o["scale"], mean, R.kmer2index)
        assert again == text, f"{tag}: record form does not reproduce the reference's text"
        p = f"e_{tag}_"
        d[p + "sha256"] = np.frombuffer(hashlib.sha256(text).digest(), dtype=np.uint8)
        d[p + "text_len"] = np.array(len(text))
        d[p + "n_lines"] = np.array(text.count(b"\n"))
        d[p + "header"] = np.frombuffer(header, dtype=np.uint8)
        d[p + "ref_to_query"] = rr.ref_to_query
        d[p + "strand"] = np.array([rr.ref_start, rr.ref_end, int(rr.is_reverse)])
        d[p + "event"] = rec["event"]
        d[p + "ref_pos"] = rec["ref_pos"]
        d[p + "label"] = rec["label"]
        d[p + "indel"] = rec["indel"]
        ap = rr.aligned_positions()
        d[p + "ap_signal"] = ap["signal"]
        d[p + "ap_core"] = ap["core"].astype(np.int32)
        d[p + "ap_residual"] = ap["residual"].astype(np.int32)
        d[p + "ap_coords"] = ap["coords"]
        d[p + "ap_ref_index"] = ap["ref_index"]
        d[p + "ap_query_index"] = ap["query_index"]
        d[p + "ap_quality"] = ap["quality"]
        print(f"{tag}: {len(text)} bytes, {d[p + 'n_lines']} lines, {rec['event'].size} records "
              f"(M {int((rec['label'] == 1).sum())}, I {int((rec['label'] == 2).sum())}), {ap['coords'].size} positions")

    # ---- builtinViterbi known answers ----
    rng = np.random.default_rng(20240611)
    obs_l, seq_l, par_l, idx_l, typ_l, score_l = [], [], [], [], [], []
    for c in range(40):
        n_bases = int(rng.integers(14, 50))
        seq = synth.BASES[rng.integers(0, 4, size=n_bases)].tobytes()
        mus = mean[synth.kmer_ranks(seq)]
        lv = []
        for m in mus:                      # skipped k-mers (deletions), stalls and junk events (insertions)
            u = rng.random()
            reps = 0 if u < 0.12 else int(rng.integers(1, 4))
            lv += [m] * reps
            if rng.random() < 0.08:
                lv.append(float(rng.normal(0, 1.5)))
        if len(lv) < 2:
            lv = [mus[0], mus[-1]]
        shift, scale, epb = float(rng.normal(90, 5)), float(rng.normal(15, 1.5)), float(rng.uniform(1.4, 2.8))
        obs = shift + scale * (np.array(lv) + 0.15 * rng.standard_normal(len(lv)))
        obs = obs.astype(np.float32).astype(np.float64)          # event means are float32-exact
        score, idx, typ = R.builtin_viterbi(obs, seq, shift, scale, epb)
        s2, i2, t2 = P.builtin_viterbi(obs, seq, shift, scale, epb, mean)
        assert score == s2 and np.array_equal(idx, i2) and np.array_equal(typ, t2), c
        obs_l.append(obs); seq_l.append(np.frombuffer(seq, dtype=np.uint8)); par_l.append([shift, scale, epb])
        idx_l.append(idx); typ_l.append(typ); score_l.append(score)
    d["v_obs"] = np.concatenate(obs_l)
    d["v_obs_off"] = np.cumsum([0] + [len(o) for o in obs_l]).astype(np.uint64)
    d["v_seq"] = np.concatenate(seq_l)
    d["v_seq_off"] = np.cumsum([0] + [len(s) for s in seq_l]).astype(np.uint64)
    d["v_par"] = np.array(par_l)
    d["v_idx"] = np.concatenate(idx_l)
    d["v_typ"] = np.concatenate(typ_l)
    d["v_path_off"] = np.cumsum([0] + [len(i) for i in idx_l]).astype(np.uint64)
    d["v_score"] = np.array(score_l)
    tt = np.concatenate(typ_l)
    print(f"viterbi: 40 windows, states D {int((tt == 0).sum())} M {int((tt == 1).sum())} I {int((tt == 2).sum())}")
    np.savez_compressed(os.path.join(OUT, "eventalign_v1.npz"), **d)


if __name__ == "__main__":
    main()
