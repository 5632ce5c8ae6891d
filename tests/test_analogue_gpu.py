"""BASELINE.json configs[3] on the device path: the resident analogue stage (dnb_batch_analogue_llr / dnb_submit_llr --
T sites, event ranges and both forward passes per site computed on the GPU from the resident alignment) against the
reference's llAcrossRead on BrdU- and EdU-substituted reads and on reads with indel CIGARs.  Tolerance 1e-4 relative
(BASELINE.json); sites and their coordinates must be identical."""
import os

import numpy as np
import pytest

from dnascent_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _table(d, ranks, name):
    t = np.zeros(4 ** 9)
    t[ranks] = d[name]
    return t


@pytest.fixture(scope="module")
def v3():
    return np.load(os.path.join(GOLDEN, "reads_v3.npz"))


def _ctx(pore_mean, unl, ana):
    c = api.Context(device=0, result_format=api.RESULT_COMPACT)
    c.load_model(api.MODEL_PORE, pore_mean)
    c.load_model(api.MODEL_UNLABELLED, *unl)
    c.load_model(api.MODEL_ANALOGUE, *ana)
    return c


def _read_and_extra(d, tag):
    q2r = d[tag + "_query_to_ref"]
    r = api.Read(None, d[tag + "_basecall"].tobytes(), d[tag + "_refseq"].tobytes(), q2r, dac=d[tag + "_dac"],
                 dac_offset=float(synth.DAC_OFFSET), dac_scale=float(synth.DAC_SCALE)).with_runs()
    x = dict(ref_to_query=d[tag + "_ref_to_query"], is_reverse=bool(int(d[tag + "_flag"]) & 16),
             ref_start=int(d[tag + "_ref_start"]), ref_end=int(d[tag + "_ref_end"]))
    return r, x


def _check(res, d, tag):
    order = np.argsort(res["pos_global"], kind="stable")          # refCoordToCalls is a std::map: ascending coordinate
    np.testing.assert_array_equal(res["pos_global"][order], d[tag + "_pos_global"], err_msg=tag)
    np.testing.assert_allclose(res["llr"][order], d[tag + "_llr"], rtol=1e-4, atol=1e-6, err_msg=tag)


def test_edu_llr_matches_reference(pore_mean, v3):
    ranks = v3["ranks"]
    c = _ctx(pore_mean, (_table(v3, ranks, "unl_mean"), _table(v3, ranks, "unl_stdv")),
             (_table(v3, ranks, "edu_mean"), _table(v3, ranks, "edu_stdv")))
    try:
        reads, extra = zip(*[_read_and_extra(v3, t) for t in ("e0", "e1")])
        b = c.submit_llr(list(reads), list(extra), 12)
        norm, res = b.results(), b.analogue_results(list(extra))
        tm = b.analogue_timings()
        b.release()
        for t, o, r in zip(("e0", "e1"), norm, res):
            np.testing.assert_array_equal(o.eventAlignment, v3[t + "_align"], err_msg=t)
            _check(r, v3, t)
        assert tm["calls"] == sum(v3[t + "_llr"].size for t in ("e0", "e1"))
    finally:
        c.close()


def test_brdu_llr_on_indel_reads_and_split_form(pore_mean, v3, golden_v2):
    ranks = v3["ranks"]
    c = _ctx(pore_mean, (_table(v3, ranks, "unl_mean"), _table(v3, ranks, "unl_stdv")),
             (_table(v3, ranks, "brdu_mean"), _table(v3, ranks, "brdu_stdv")))
    try:
        reads, extra = zip(*[_read_and_extra(v3, t) for t in ("j0", "j1")])
        b = c.upload(list(reads))                   # split form: upload / run / analogue stage on the resident batch
        b.run()
        res = b.analogue_llr(list(extra), 12)
        b.release()
        for t, r in zip(("j0", "j1"), res):
            _check(r, v3, t)
    finally:
        c.close()
    # the BrdU-substituted `{L}M` reads of reads_v2 (forward and reverse strand) through the same stage
    g2, (um, us, am, as_), _ = golden_v2
    c = _ctx(pore_mean, (um, us), (am, as_))
    try:
        rs, xs = [], []
        for t in ("a0", "a1"):
            g = g2[t]
            rs.append(api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                               dac_scale=float(synth.DAC_SCALE)))
            rev = bool(g.flag & 16)
            xs.append(dict(ref_to_query=np.arange(len(g.refseq), dtype=np.int32), is_reverse=rev, ref_start=g.pos,
                           ref_end=g.pos + len(g.refseq)))
        b = c.submit_llr(rs, xs, 12)
        res = b.analogue_results(xs)
        b.release()
        for t, r in zip(("a0", "a1"), res):
            order = np.argsort(r["pos_global"], kind="stable")
            np.testing.assert_array_equal(r["pos_global"][order], g2[t].pos_global, err_msg=t)
            np.testing.assert_allclose(r["llr"][order], g2[t].llr, rtol=1e-4, atol=1e-6, err_msg=t)
    finally:
        c.close()


def test_flat_sequence_probability_matches_resident_stage(pore_mean, v3):
    """dnb_sequence_probability_batch (observations given by the caller) and the resident stage share the forward-pass
    code: the host-side gathering of api.gather_sites + the flat entry must give the same numbers as the device path."""
    ranks = v3["ranks"]
    c = _ctx(pore_mean, (_table(v3, ranks, "unl_mean"), _table(v3, ranks, "unl_stdv")),
             (_table(v3, ranks, "brdu_mean"), _table(v3, ranks, "brdu_stdv")))
    try:
        r, x = _read_and_extra(v3, "j0")
        b = c.submit_llr([r], [x], 12)
        o, res = b.results()[0], b.analogue_results([x])[0]
        b.release()
        sites = api.gather_sites(r.referenceSeqMappedTo, x["ref_to_query"], x["is_reverse"], o.eventAlignment, o.event_mean, 12)
        assert [s[0] for s in sites] == res["pos_on_ref"].tolist()
        la, lt = c.sequence_probability_batch([s[1] for s in sites], [s[2] for s in sites], o.shift, o.scale,
                                              o.eventsPerBase, 12)
        np.testing.assert_allclose(la, res["log_analogue"], rtol=1e-12)
        np.testing.assert_allclose(lt, res["log_thymidine"], rtol=1e-12)
    finally:
        c.close()
