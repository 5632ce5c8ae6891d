"""SURVEY s.8 row f1 on the GPU: dnb_eventalign_batch (dnascent_b200/csrc/eventalign.cu), called through the C ABI,
against the reference's own output (tests/golden/eventalign_v1.npz: records and the sha256 of the text the reference
printed) and against the CPU oracle port on seeded inputs.  The state path is a discrete result: identical records
are required (the emission log-densities may differ from glibc's in the last ulps, see eventalign.cu's header)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers.eventalign_render import render
from test_eventalign_cpu import all_golden_reads, golden_eventalign_inputs
from dnascent_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ea_golden():
    return np.load(os.path.join(GOLDEN, "eventalign_v1.npz"))


def test_eventalign_golden_records_and_text(ctx, pore_mean, ea_golden, golden_reads, golden_v2):
    reads = all_golden_reads(golden_reads, golden_v2)
    out = ctx.eventalign([golden_eventalign_inputs(g, ea_golden, tag) for tag, g in reads], window=50)
    for (tag, g), rec in zip(reads, out):
        p = f"e_{tag}_"
        assert rec["status"] == api.READ_OK, tag
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(rec[key], ea_golden[p + key], err_msg=f"{tag} {key}")
        ref_start, ref_end, is_rev = (int(x) for x in ea_golden[p + "strand"])
        text = render(ea_golden[p + "header"].tobytes(), g.refseq, ref_start, ref_end, bool(is_rev),
                      g.raw.astype(np.float64)[:int(g.event_raw_len.sum())], g.event_raw_len, rec, g.shift, g.scale,
                      pore_mean, lambda km: int(synth.kmer_ranks(km)[0]))
        assert hashlib.sha256(text).digest() == ea_golden[p + "sha256"].tobytes(), tag


def test_eventalign_viterbi_windows(ctx, port, pore_mean, ea_golden):
    """builtinViterbi known answers: each hand-made window becomes a one-window read (identity refToQuery, every
    event aligned inside it); the insertion / deletion states of the reference's path must come back."""
    e = ea_golden
    reads, expect = [], []
    for c in range(e["v_score"].size):
        obs = e["v_obs"][int(e["v_obs_off"][c]):int(e["v_obs_off"][c + 1])]
        seq = e["v_seq"][int(e["v_seq_off"][c]):int(e["v_seq_off"][c + 1])].tobytes()
        shift, scale, epb = e["v_par"][c]
        n = len(seq) - 8
        al = np.stack([np.arange(obs.size), np.minimum(np.arange(obs.size) * n // obs.size, n - 1)], axis=1).astype(np.uint32)
        r = dict(refseq=seq, ref_to_query=np.arange(len(seq), dtype=np.int32), eventAlignment=al,
                 event_mean=obs.astype(np.float32), shift=shift, scale=scale, events_per_base=epb)
        reads.append(r)
        expect.append(port.eventalign(seq, r["ref_to_query"], al[:, 0], al[:, 1], obs, shift, scale, epb, pore_mean))
        # first window of the chain == the reference's path (emitting states up to the last match)
        lo, hi = int(e["v_path_off"][c]), int(e["v_path_off"][c + 1])
        idx, typ = e["v_idx"][lo:hi], e["v_typ"][lo:hi]
        emit = typ != 0
        last_m = np.nonzero(typ[emit] == 1)[0]
        if last_m.size:
            k = int(last_m[-1]) + 1
            np.testing.assert_array_equal(expect[-1]["ref_pos"][:k], idx[emit][:k])
            np.testing.assert_array_equal(expect[-1]["label"][:k], typ[emit][:k])
    out = ctx.eventalign(reads, window=50)
    n_i = 0
    for c, (rec, exp) in enumerate(zip(out, expect)):
        assert rec["status"] == api.READ_OK, c
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(rec[key], exp[key], err_msg=f"window {c} {key}")
        n_i += int((rec["label"] == 2).sum())
    assert n_i > 0


def test_eventalign_after_normalise_vs_port(ctx, port, pore_mean):
    """The stage as the read loop runs it (detect.cpp:876-888): the product's normaliseEvents output fed to the
    product's eventalign, against the port fed the same arrays; forward and reverse reads, with substitutions."""
    ref = synth.make_reference(200_000, 91)
    base = synth.simulate_batch(ref, [9000, 12000, 7000, 15000, 10000, 8000, 30000, 2000], pore_mean, seed=92, sub_rate=0.01)
    res = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True) for r in base])
    reads, keep = [], []
    for sr, o in zip(base, res):
        if o.status != api.READ_OK:
            continue
        r2q = np.zeros(len(sr.refseq), dtype=np.int32)
        q2r = np.asarray(sr.query_to_ref)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]          # plain {L}M reads: refToQuery is the inverse map
        reads.append(dict(refseq=sr.refseq, ref_to_query=r2q, eventAlignment=o.eventAlignment, event_mean=o.event_mean,
                          shift=o.shift, scale=o.scale, events_per_base=o.eventsPerBase))
        keep.append(o)
    assert len(reads) >= 6
    out = ctx.eventalign(reads, window=50)
    for r, rec in zip(reads, out):
        exp = port.eventalign(r["refseq"], r["ref_to_query"], r["eventAlignment"][:, 0], r["eventAlignment"][:, 1],
                              r["event_mean"].astype(np.float64), r["shift"], r["scale"], r["events_per_base"], pore_mean)
        assert rec["status"] == api.READ_OK
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(rec[key], exp[key], err_msg=key)
        assert rec["event"].size > 0.5 * r["eventAlignment"].shape[0]


def test_eventalign_edge_cases(ctx, pore_mean):
    seq = synth.make_reference(200, 5)
    base = dict(refseq=seq, ref_to_query=np.arange(200, dtype=np.int32), eventAlignment=np.zeros((0, 2), dtype=np.uint32),
                event_mean=np.zeros(0, dtype=np.float32), shift=90.0, scale=15.0, events_per_base=2.0)
    undefined = dict(base, events_per_base=1.0)                              # eln(0) -> NaN -> NegativeLog in the reference
    with_n = dict(base, refseq=seq[:60] + b"N" * 30 + seq[90:],              # windows with an undefined base are skipped
                  eventAlignment=np.stack([np.arange(300), np.arange(300) * 192 // 300], axis=1).astype(np.uint32),
                  event_mean=np.full(300, 95.0, dtype=np.float32))
    out = ctx.eventalign([base, undefined, with_n], window=50)
    assert out[0]["status"] == api.READ_OK and out[0]["event"].size == 0    # no events: header only
    assert out[1]["status"] == api.READ_UNDEFINED
    assert out[2]["status"] == api.READ_OK
    assert not np.any((out[2]["ref_pos"] >= 52) & (out[2]["ref_pos"] < 90))
    assert ctx.eventalign([], window=50) == []


def test_eventalign_widest_window_uses_third_register_slot(ctx, port, pore_mean):
    """The breakpoint search (alignment.cpp:566-592) can stretch a window to 73 bases = 65 HMM states, the only case in
    which the kernel's third register slot (states 64..) is live.  These two reference prefixes have their first
    qualifying breakpoint exactly at i = 64 (found by brute force with the rule restated below); reads starting on
    them must come back identical to the oracle's, forward and reverse-complement strand handling included."""
    prefixes = [b"GGATGAGCGGTTACCTCCTTGCCGCTGGCAGTCTTTTTAAACTGCCTTGATATTGCAAGGATGGGACTTTATAGG",
                b"TACGTGGTAACAGAGTATTTCCATCAAAATACGTTTAGATGCTGGTTGTAACGAAAAGGATTGCACAGCCGCTCT"]

    def first_break(seq, W=50, k=9):
        m = pore_mean[synth.kmer_ranks(seq[:int(1.5 * W)])]
        for i in range(W, int(1.5 * W) - k - 1):
            if abs(m[i] - m[i + 1]) > 0.75 and abs(m[i] - m[i - 1]) > 0.75:
                return i
        return None

    rng = np.random.default_rng(77)
    base = []
    for j, pre in enumerate(prefixes):
        assert first_break(pre) == 64
        ref = pre + synth.make_reference(6000, 300 + j)
        base.append(synth.simulate_read(ref, 0, 5000, False, pore_mean, rng, name=f"w{j}"))
    res = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True) for r in base])
    reads = []
    for sr, o in zip(base, res):
        assert o.status == api.READ_OK
        assert sr.refseq[:75] in prefixes
        reads.append(dict(refseq=sr.refseq, ref_to_query=np.arange(len(sr.refseq), dtype=np.int32),
                          eventAlignment=o.eventAlignment, event_mean=o.event_mean, shift=o.shift, scale=o.scale,
                          events_per_base=o.eventsPerBase))
    out = ctx.eventalign(reads, window=50)
    for r, rec in zip(reads, out):
        exp = port.eventalign(r["refseq"], r["ref_to_query"], r["eventAlignment"][:, 0], r["eventAlignment"][:, 1],
                              r["event_mean"].astype(np.float64), r["shift"], r["scale"], r["events_per_base"], pore_mean)
        assert rec["status"] == api.READ_OK
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(rec[key], exp[key], err_msg=key)
        assert rec["event"].size > 0.5 * r["eventAlignment"].shape[0]
