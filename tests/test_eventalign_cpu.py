"""SURVEY s.8 row f1 on the CPU: the oracle port of builtinViterbi / eventalign (oracle/dnb_oracle.c) against the
fixtures produced by the unmodified reference (tests/golden/eventalign_v1.npz) and, where oracle/_ref is built,
against the reference itself on fresh seeded reads (byte-identical text)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers.eventalign_render import render
from dnascent_b200 import synth


@pytest.fixture(scope="module")
def ea_golden():
    return np.load(os.path.join(GOLDEN, "eventalign_v1.npz"))


def golden_eventalign_inputs(g, e, tag):
    p = f"e_{tag}_"
    return dict(refseq=g.refseq, ref_to_query=e[p + "ref_to_query"], eventAlignment=g.align, event_mean=g.event_mean,
                shift=g.shift, scale=g.scale, events_per_base=g.events_per_base)


def all_golden_reads(golden_reads, golden_v2):
    out = [(f"g{g.index}", g) for g in golden_reads]
    out += [(t, golden_v2[0][t]) for t in ("a0", "a1", "i0", "i1")]
    return out


def test_port_viterbi_matches_reference_paths(port, pore_mean, ea_golden):
    e = ea_golden
    for c in range(e["v_score"].size):
        obs = e["v_obs"][int(e["v_obs_off"][c]):int(e["v_obs_off"][c + 1])]
        seq = e["v_seq"][int(e["v_seq_off"][c]):int(e["v_seq_off"][c + 1])].tobytes()
        shift, scale, epb = e["v_par"][c]
        score, idx, typ = port.builtin_viterbi(obs, seq, shift, scale, epb, pore_mean)
        lo, hi = int(e["v_path_off"][c]), int(e["v_path_off"][c + 1])
        assert score == e["v_score"][c]
        np.testing.assert_array_equal(idx, e["v_idx"][lo:hi])
        np.testing.assert_array_equal(typ, e["v_typ"][lo:hi])


def test_port_eventalign_matches_reference_records(port, pore_mean, ea_golden, golden_reads, golden_v2):
    for tag, g in all_golden_reads(golden_reads, golden_v2):
        inp = golden_eventalign_inputs(g, ea_golden, tag)
        rec = port.eventalign(inp["refseq"], inp["ref_to_query"], g.align[:, 0], g.align[:, 1],
                              g.event_mean.astype(np.float64), g.shift, g.scale, g.events_per_base, pore_mean)
        p = f"e_{tag}_"
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(rec[key], ea_golden[p + key], err_msg=f"{tag} {key}")
        # and the text: rebuilt from the records it must hash to the reference's output
        raw_len = g.event_raw_len
        ref_start, ref_end, is_rev = (int(x) for x in ea_golden[p + "strand"])
        # r.events[j].raw are consecutive slices of the raw signal starting at sample 0 (event_handling.cpp:549-575)
        text = render(ea_golden[p + "header"].tobytes(), g.refseq, ref_start, ref_end, bool(is_rev),
                      g.raw.astype(np.float64)[:int(raw_len.sum())], raw_len, rec, g.shift, g.scale, pore_mean,
                      lambda km: int(synth.kmer_ranks(km)[0]))
        assert len(text) == int(ea_golden[p + "text_len"]), tag
        assert hashlib.sha256(text).digest() == ea_golden[p + "sha256"].tobytes(), tag


def golden_records(e, tag):
    p = f"e_{tag}_"
    return {key: e[p + key] for key in ("event", "ref_pos", "label", "indel")}


def event_starts(raw_len):
    return np.concatenate([[0], np.cumsum(raw_len)]).astype(np.uint32)


AP_KEYS = ("signal", "core", "residual", "coords", "ref_index", "query_index", "quality")


def test_port_dnn_features_match_reference_tensors(port, ea_golden, golden_reads, golden_v2):
    """Row f2: the port's tensor builder on the golden records == makeSignalTensor / makeCoreSequenceTensor /
    makeResidualSequenceTensor / getReferenceCoords / ...Indices of the unmodified reference (reads.h:305-427)."""
    e = ea_golden
    for tag, g in all_golden_reads(golden_reads, golden_v2):
        p = f"e_{tag}_"
        ref_start, ref_end, is_rev = (int(x) for x in e[p + "strand"])
        f = port.dnn_features(g.refseq, e[p + "ref_to_query"], is_rev, ref_start, ref_end, golden_records(e, tag),
                              g.raw.astype(np.float64), event_starts(g.event_raw_len), g.shift, g.scale)
        for key in AP_KEYS:
            np.testing.assert_array_equal(f[key], e[p + "ap_" + key], err_msg=f"{tag} {key}")
        # positions already called (refCoordToCalls, alignment.cpp:711) are not added
        called = np.sort(e[p + "ap_coords"][::3])
        f2 = port.dnn_features(g.refseq, e[p + "ref_to_query"], is_rev, ref_start, ref_end, golden_records(e, tag),
                               g.raw.astype(np.float64), event_starts(g.event_raw_len), g.shift, g.scale, called=called)
        keep = ~np.isin(e[p + "ap_coords"], called)
        for key in AP_KEYS:
            np.testing.assert_array_equal(f2[key], e[p + "ap_" + key][keep], err_msg=f"{tag} {key} (called)")


def test_port_eventalign_vs_reference_fresh_reads(port, ref_oracle, pore_mean):
    ref = synth.make_reference(80_000, 77)
    ref_oracle.set_reference(ref)
    reads = synth.simulate_batch(ref, [2500, 4000, 3200, 5200], pore_mean, seed=78, sub_rate=0.01)
    for sr in reads:
        h = ref_oracle.read_new(sr)
        o = h.normalise(staged=False)
        if o["align_event"].size == 0:
            continue
        text = h.eventalign(50)
        rec = port.eventalign(h.refseq, h.ref_to_query, o["align_event"], o["align_kmer"], o["event_mean"], o["shift"],
                              o["scale"], o["events_per_base"], pore_mean)
        mine = render(text.split(b"\n")[0] + b"\n", h.refseq, h.ref_start, h.ref_end, h.is_reverse,
                      h.events_raw_concat(), o["event_raw_len"], rec, o["shift"], o["scale"], pore_mean,
                      ref_oracle.kmer2index)
        assert mine == text
        ap = h.aligned_positions()
        f = port.dnn_features(h.refseq, h.ref_to_query, h.is_reverse, h.ref_start, h.ref_end, rec, h.events_raw_concat(),
                              event_starts(o["event_raw_len"]), o["shift"], o["scale"])
        for key in AP_KEYS:
            np.testing.assert_array_equal(f[key], ap[key], err_msg=key)


def test_window_parallel_eventalign_design_check(port, pore_mean, ea_golden, golden_reads, golden_v2):
    """DESIGN.md s.8 item 2 (tests/helpers/proto_window_parallel_eventalign.py): eventalign with all windows of a read run
    independently on the 'every window advances fully' chain, plus a verify-and-repair walk, gives exactly the records the
    UNMODIFIED REFERENCE produced for the golden reads (indel / soft-clip CIGARs and analogue reads included), and the repair
    rounds touch only a few windows."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "proto_window_parallel_eventalign.py")
    spec = importlib.util.spec_from_file_location("proto_window_parallel_eventalign", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    total0 = repair = most_rounds = 0
    for tag, g in all_golden_reads(golden_reads, golden_v2):
        serial = golden_records(ea_golden, tag)
        R = mod.Read(port, pore_mean, g.refseq, ea_golden[f"e_{tag}_ref_to_query"], g.align[:, 0], g.align[:, 1],
                     g.event_mean.astype(np.float64), g.shift, g.scale, g.events_per_base, serial=serial)
        rounds = mod.check(R)
        total0 += rounds[0]
        repair += sum(rounds[1:])
        most_rounds = max(most_rounds, len(rounds))
    # ~5 % of the windows are re-run (mostly on the read with a long insertion/deletion CIGAR), in one extra round
    assert total0 > 300 and repair < 0.15 * total0 and most_rounds <= 3
