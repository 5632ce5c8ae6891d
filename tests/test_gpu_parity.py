"""GPU parity tests: the CUDA path, called through the C ABI, against the golden fixtures (outputs of the unmodified
reference) and against the CPU oracle on seeded inputs.  Bit-exact for event boundaries, event means, alignment
paths, QC flags and scalings (integer / IEEE-exact work); 1e-4 relative for the analogue log-likelihoods."""
import numpy as np
import pytest

from dnascent_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check_against_golden(res, g):
    assert res.status == api.READ_OK
    assert res.et_n == g.et_mean.size
    np.testing.assert_array_equal(res.event_mean, g.event_mean)
    np.testing.assert_array_equal(np.diff(res.event_start.astype(np.int64)), g.event_raw_len.astype(np.int64))
    np.testing.assert_array_equal(res.eventAlignment, g.align)
    assert res.rough_shift == g.rough_shift and res.rough_scale == g.rough_scale
    assert res.shift == g.shift and res.scale == g.scale
    assert res.eventsPerBase == g.events_per_base
    assert res.avg_log_emission == g.avg_log_emission
    assert res.spanned == g.spanned and res.maxGap == g.max_gap
    np.testing.assert_array_equal(res.cleaned_signal, g.cleaned_signal)
    np.testing.assert_array_equal(res.cleaned_rank, g.cleaned_rank)


def test_golden_reads_float_input(ctx, golden_reads):
    out = ctx.normaliseEvents([api.Read(g.raw, g.basecall, g.refseq, g.query_to_ref) for g in golden_reads])
    for res, g in zip(out, golden_reads):
        _check_against_golden(res, g)


def test_golden_reads_int16_input(ctx, golden_reads):
    reads = [api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                      dac_scale=float(synth.DAC_SCALE)) for g in golden_reads]
    for res, g in zip(ctx.normaliseEvents(reads), golden_reads):
        _check_against_golden(res, g)


def test_detect_events_drop_in(ctx, golden_reads):
    for g in golden_reads[:3]:
        st, ln, mn, sd = ctx.detect_events(g.raw)
        np.testing.assert_array_equal(st, g.et_start.astype(np.uint64))
        np.testing.assert_array_equal(ln, g.et_length)
        np.testing.assert_array_equal(mn, g.et_mean)
        np.testing.assert_array_equal(sd, g.et_stdv)
