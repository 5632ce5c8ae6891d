"""Host-side mirror of the reference interface for the detect signal hot path, over the C ABI.

The reference's interface for this path is one C++ call per read, ``normaliseEvents(DNAscent::read&, bool)``
(/root/reference/src/event_handling.h:13), plus ``detect_events`` (src/scrappie/event_detection.h:35), the
log-space helpers of src/probability.h:26-33 and ``sequenceProbability`` (src/detect.h:119).  The names, argument
meaning and failure convention (empty ``eventAlignment`` == failed read, src/detect.cpp:879) are kept; the call is
batched because a GPU wants many reads per launch.  PyTorch is not involved: the library owns its device memory.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import numpy as np

from . import _lib
from ._lib import Config, ReadDesc, ReadResult, EventT, DnbError

MODEL_PORE, MODEL_UNLABELLED, MODEL_ANALOGUE = 0, 1, 2
RESULT_DENSE, RESULT_COMPACT = 0, 1
READ_OK, READ_QC_FAIL, READ_SCALE_FAIL, READ_UNDEFINED, READ_OVERFLOW = 0, 1, 2, 3, 4
ERR_NEGATIVE_LOG = 6


class NegativeLog(ValueError):
    """src/probability.h:11-17"""

    def __str__(self):
        return "Negative value passed to natural log function."


@dataclasses.dataclass
class Read:
    """The fields of DNAscent::read that normaliseEvents consumes (src/reads.h:178-208)."""
    raw: np.ndarray | None          # float32 pA (float32-exact, src/pod5.cpp:60) ...
    basecall: bytes
    referenceSeqMappedTo: bytes
    queryToRef: np.ndarray          # dense int32, -1 = no entry
    dac: np.ndarray | None = None   # ... or int16 DAC with calibration
    dac_offset: float = 0.0
    dac_scale: float = 1.0
    q2r_runs: np.ndarray | None = None   # queryToRef as runs (dtype _lib.Q2R_RUN_DTYPE) instead of the dense array

    def with_runs(self):
        """The same read with queryToRef given as runs (16 B per CIGAR operation over PCIe instead of 4 B per base)."""
        return dataclasses.replace(self, queryToRef=None, q2r_runs=q2r_to_runs(self.queryToRef))

    @classmethod
    def from_synth(cls, sr, use_dac: bool = False):
        from .synth import DAC_OFFSET, DAC_SCALE
        if use_dac:
            return cls(None, sr.basecall, sr.refseq, sr.query_to_ref, dac=sr.dac, dac_offset=float(DAC_OFFSET),
                       dac_scale=float(DAC_SCALE))
        return cls(sr.raw, sr.basecall, sr.refseq, sr.query_to_ref)


@dataclasses.dataclass
class Normalised:
    """What normaliseEvents leaves in DNAscent::read: events, eventAlignment, scalings, alignmentQCs."""
    status: int
    et_n: int
    event_start: np.ndarray     # uint32 [n_events+1]; events[j].raw == raw[event_start[j]:event_start[j+1]]
    event_mean: np.ndarray      # float32 [n_events]  events[j].mean
    eventAlignment: np.ndarray  # uint32 [n_align, 2] (event_idx, kmer_idx); empty == failed read
    shift: float
    scale: float
    eventsPerBase: float
    rough_shift: float
    rough_scale: float
    avg_log_emission: float
    spanned: bool
    maxGap: int
    cleaned_signal: np.ndarray | None = None
    cleaned_rank: np.ndarray | None = None


def q2r_to_runs(q2r) -> np.ndarray:
    """Dense queryToRef (int32, -1 = no entry) -> dnb_q2r_run array: maximal runs with stride 1 (aligned bases) or
    stride 0 (the constant entries parseCigar gives insertions / soft clips, src/htsInterface.cpp:143-152)."""
    q2r = np.asarray(q2r, dtype=np.int64)
    n = q2r.size
    runs = []
    i = 0
    while i < n:
        if q2r[i] < 0:
            i += 1
            continue
        j = i + 1
        stride = 1
        if j < n and q2r[j] == q2r[i]:
            stride = 0
        while j < n and q2r[j] >= 0 and q2r[j] == q2r[i] + stride * (j - i):
            j += 1
        runs.append((i, j - i, int(q2r[i]), stride))
        i = j
    return np.array(runs, dtype=_lib.Q2R_RUN_DTYPE) if runs else np.zeros(0, dtype=_lib.Q2R_RUN_DTYPE)


def _as(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _extras(extra):
    """dnb_read_extra array for a list of dicts (ref_to_query, is_reverse, ref_start, ref_end[, called])."""
    n = len(extra)
    ex = (_lib.ReadExtra * max(n, 1))()
    keep = []
    for i, r in enumerate(extra):
        r2q = np.ascontiguousarray(r["ref_to_query"], dtype=np.int32)
        called = np.ascontiguousarray(r.get("called", []), dtype=np.uint32)
        keep += [r2q, called]
        ex[i].ref_to_query = r2q.ctypes.data
        ex[i].is_reverse, ex[i].ref_start, ex[i].ref_end = int(bool(r["is_reverse"])), int(r["ref_start"]), int(r["ref_end"])
        ex[i].called, ex[i].n_called = (called.ctypes.data if called.size else None), called.size
    return ex, keep


class Batch:
    def __init__(self, ctx: "Context", handle, n_reads: int, keepalive):
        self.ctx, self.h, self.n, self._keep = ctx, handle, n_reads, keepalive

    def wait(self):
        _lib.check(self.ctx.L.dnb_wait(self.h), "dnb_wait")

    def run(self):
        _lib.check(self.ctx.L.dnb_batch_run(self.h), "dnb_batch_run")

    def fetch(self):
        _lib.check(self.ctx.L.dnb_batch_fetch(self.h), "dnb_batch_fetch")

    def drop_workspace(self):
        _lib.check(self.ctx.L.dnb_batch_drop_workspace(self.h), "dnb_batch_drop_workspace")

    def timings(self):
        ms = (C.c_double * 8)()
        cnt = (C.c_uint64 * 8)()
        _lib.check(self.ctx.L.dnb_batch_timings(self.h, C.byref(ms), C.byref(cnt)), "dnb_batch_timings")
        names = ("segmentation", "prep", "banded_dp", "backtrace", "theil_sen", "total", "host_step", "wall")
        cn = ("samples", "events", "kmers", "bands", "cells", "launches", "seg_serial_reads", "failed_reads")
        out = dict(zip(names, ms))
        seg = (C.c_double * 3)()
        if self.ctx.L.dnb_batch_seg_timings(self.h, C.byref(seg)) == 0:
            out.update(seg_checkpoint=seg[0], seg_tiles=seg[1], seg_rest=seg[2])
        return out, dict(zip(cn, (int(x) for x in cnt)))

    def device(self) -> int:
        return int(self.ctx.L.dnb_batch_device(self.h))

    def io_bytes(self):
        """(host->device, device->host) bytes this batch moved over PCIe, counted by the library."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _lib.check(self.ctx.L.dnb_batch_io_bytes(self.h, C.byref(a), C.byref(b)), "dnb_batch_io_bytes")
        return a.value, b.value

    def result(self, i: int) -> Normalised:
        r = ReadResult()
        _lib.check(self.ctx.L.dnb_result(self.h, i, C.byref(r)), "dnb_result")
        ne = r.n_events
        if self.ctx.cfg.result_format == RESULT_COMPACT:
            # compact wire format: rebuild the dense arrays with the library's own expanders
            starts = np.zeros(ne + 1 if ne else 0, dtype=np.uint32)
            pairs = np.zeros((r.n_align, 2), dtype=np.uint32)
            _lib.check(self.ctx.L.dnb_expand_events(C.byref(r), starts.ctypes.data), "dnb_expand_events")
            _lib.check(self.ctx.L.dnb_expand_alignment(C.byref(r), pairs.ctypes.data), "dnb_expand_alignment")
        else:
            starts = _as(r.event_start, ne + 1 if ne else 0, np.uint32)
            pairs = _as(r.align_pairs, 2 * r.n_align, np.uint32).reshape(-1, 2)
        return Normalised(
            status=r.status, et_n=r.et_n, event_start=starts,
            event_mean=_as(r.event_mean, ne, np.float32), eventAlignment=pairs, shift=r.shift, scale=r.scale,
            eventsPerBase=r.events_per_base, rough_shift=r.rough_shift, rough_scale=r.rough_scale,
            avg_log_emission=r.avg_log_emission, spanned=bool(r.spanned), maxGap=r.max_gap,
            cleaned_signal=_as(r.cleaned_signal, r.n_cleaned, np.float64) if r.n_cleaned else None,
            cleaned_rank=_as(r.cleaned_rank, r.n_cleaned, np.uint32) if r.n_cleaned else None)

    def results(self):
        return [self.result(i) for i in range(self.n)]

    # -- resident eventalign + DNN input tensors (detect.cpp:888 + runCNN's inputs) on what run() left in HBM --------
    def eventalign_features(self, extra, window: int = 50, want_records: bool = False):
        """`extra`: per read a dict with ref_to_query (int32[ref_len]), is_reverse, ref_start, ref_end and optionally
        called (sorted uint32).  Returns per read the dict Context.eventalign_features returns."""
        assert len(extra) == self.n
        ex, keep = _extras(extra)
        _lib.check(self.ctx.L.dnb_batch_eventalign_features(self.h, C.addressof(ex), window, int(want_records)),
                   "dnb_batch_eventalign_features")
        return self.feature_results(want_records)

    def feature_results(self, want_records: bool = False):
        out = []
        for i in range(self.n):
            fr = _lib.FeatureResult()
            _lib.check(self.ctx.L.dnb_batch_feature_result(self.h, i, C.byref(fr)), "dnb_batch_feature_result")
            P = fr.n_pos
            o = dict(status=fr.status, signal=_as(fr.signal, P * _lib.RAWDEPTH, np.float32).reshape(P, _lib.RAWDEPTH),
                     core=_as(fr.core, P, np.float32), residual=_as(fr.residual, P, np.float32),
                     coords=_as(fr.coords, P, np.uint32), ref_index=_as(fr.ref_index, P, np.uint32),
                     query_index=_as(fr.query_index, P, np.uint32), quality=_as(fr.quality, P, np.int32))
            if want_records:
                n = fr.n_recs
                rr = (np.frombuffer(C.string_at(fr.recs, n * _lib.EVENTALIGN_REC_DTYPE.itemsize), dtype=_lib.EVENTALIGN_REC_DTYPE)
                      if n else np.zeros(0, dtype=_lib.EVENTALIGN_REC_DTYPE))
                o.update(event=rr["event"].copy(), ref_pos=rr["ref_pos"].copy(), label=rr["label"].astype(np.uint8),
                         indel=rr["indel_score"].copy())
            out.append(o)
        return out

    # -- resident analogue stage: llAcrossRead (detect.cpp:393-574) on what run() left in HBM -------------------------
    def analogue_llr(self, extra, window: int = 12):
        """`extra`: per read a dict with ref_to_query (int32[ref_len]) and is_reverse (ref_start / ref_end are only used
        for the global coordinate returned here).  Returns analogue_results()."""
        assert len(extra) == self.n
        ex, keep = _extras(extra)
        _lib.check(self.ctx.L.dnb_batch_analogue_llr(self.h, C.addressof(ex), window), "dnb_batch_analogue_llr")
        return self.analogue_results(extra)

    def analogue_results(self, extra=None):
        """Per read a dict: pos_on_ref, n_events, log_analogue, log_thymidine, llr of the sites where a call was made,
        in llAcrossRead's visiting order; with `extra` also pos_global as the reference computes it (:530-538)."""
        out = []
        for i in range(self.n):
            ar = _lib.AnalogueResult()
            _lib.check(self.ctx.L.dnb_batch_analogue_result(self.h, i, C.byref(ar)), "dnb_batch_analogue_result")
            n = ar.n_sites
            nev = _as(ar.n_events, n, np.uint32)
            keep = nev > 0
            pos = _as(ar.pos_on_ref, n, np.uint32)[keep]
            la, lt = _as(ar.log_analogue, n, np.float64)[keep], _as(ar.log_thymidine, n, np.float64)[keep]
            o = dict(status=ar.status, n_sites=int(n), pos_on_ref=pos, n_events=nev[keep], log_analogue=la, log_thymidine=lt,
                     llr=la - lt)
            if extra is not None:
                x = extra[i]
                o["pos_global"] = (int(x["ref_end"]) - pos.astype(np.int64) - 1 if x["is_reverse"]
                                   else int(x["ref_start"]) + pos.astype(np.int64))
            out.append(o)
        return out

    def analogue_timings(self):
        ms = (C.c_double * 2)()
        cnt = (C.c_uint64 * 4)()
        _lib.check(self.ctx.L.dnb_batch_analogue_timings(self.h, C.byref(ms), C.byref(cnt)), "dnb_batch_analogue_timings")
        return dict(sites_kernel_ms=ms[0], forward_kernel_ms=ms[1], candidate_sites=int(cnt[0]), calls=int(cnt[1]),
                    observations=int(cnt[2]), d2h_bytes=int(cnt[3]))

    def feature_rows(self) -> int:
        """Total tensor rows of the batch (sum of dnb_feature_result.n_pos), without copying the tensors."""
        fr = _lib.FeatureResult()
        tot = 0
        for i in range(self.n):
            _lib.check(self.ctx.L.dnb_batch_feature_result(self.h, i, C.byref(fr)), "dnb_batch_feature_result")
            tot += fr.n_pos
        return tot

    def stage2_timings(self):
        ms = (C.c_double * 2)()
        by = (C.c_uint64 * 2)()
        _lib.check(self.ctx.L.dnb_batch_stage2_timings(self.h, C.byref(ms), C.byref(by)), "dnb_batch_stage2_timings")
        return dict(eventalign_kernel_ms=ms[0], features_kernel_ms=ms[1], h2d_bytes=int(by[0]), d2h_bytes=int(by[1]))

    def release(self):
        if self.h:
            self.ctx.L.dnb_release(self.h)
            self.h = None
            self._keep = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Context:
    """One per process and GPU (``dnb_ctx``).  Fails loudly without a CUDA device: there is no CPU path."""

    def __init__(self, device: int = 0, keep_debug: bool = False, event_capacity_per_sample: float | None = None,
                 result_format: int = RESULT_DENSE, devices=None, workspace_bytes: int = 0):
        self.L = _lib.lib()
        cfg = Config()
        self.L.dnb_default_config(C.byref(cfg))
        cfg.device = device
        cfg.keep_debug = int(keep_debug)
        cfg.result_format = result_format
        cfg.workspace_bytes = workspace_bytes
        if devices is not None:          # one context driving several GPUs (dnb_config.devices)
            cfg.n_devices = len(devices)
            for k, dv in enumerate(devices):
                cfg.devices[k] = int(dv)
        if event_capacity_per_sample is not None:
            cfg.event_capacity_per_sample = event_capacity_per_sample
        self.cfg = cfg
        h = C.c_void_p()
        _lib.check(self.L.dnb_create(C.byref(h), C.byref(cfg)), "dnb_create")
        self.h = h

    def close(self):
        if self.h:
            self.L.dnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """dnb_trim: give the idle device / pinned blocks the context keeps cached back to the driver."""
        _lib.check(self.L.dnb_trim(self.h), "dnb_trim")

    def load_model(self, which: int, mean: np.ndarray, stdv: np.ndarray | None = None):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        sd = None if stdv is None else np.ascontiguousarray(stdv, dtype=np.float64)
        _lib.check(self.L.dnb_load_model(self.h, which, mean.ctypes.data, None if sd is None else sd.ctypes.data,
                                         mean.size), "dnb_load_model")

    # -- batches ------------------------------------------------------------------------------------
    @staticmethod
    def _descs(reads):
        n = len(reads)
        arr = (ReadDesc * n)()
        keep = []
        for i, r in enumerate(reads):
            d = arr[i]
            if r.dac is not None:
                dac = np.ascontiguousarray(r.dac, dtype=np.int16)
                keep.append(dac)
                d.raw_dac, d.raw_pA, d.n_samples = dac.ctypes.data, None, dac.size
                d.dac_offset, d.dac_scale = r.dac_offset, r.dac_scale
            else:
                raw = np.ascontiguousarray(r.raw, dtype=np.float32)
                keep.append(raw)
                d.raw_pA, d.raw_dac, d.n_samples = raw.ctypes.data, None, raw.size
            d.query, d.query_len = r.basecall, len(r.basecall)
            d.ref, d.ref_len = r.referenceSeqMappedTo, len(r.referenceSeqMappedTo)
            if r.q2r_runs is not None:
                runs = np.ascontiguousarray(r.q2r_runs, dtype=_lib.Q2R_RUN_DTYPE)
                if runs.size == 0:      # a non-NULL pointer says "runs given", even when there are none
                    runs = np.zeros(1, dtype=_lib.Q2R_RUN_DTYPE)
                    d.n_q2r_runs = 0
                else:
                    d.n_q2r_runs = runs.size
                keep.append(runs)
                d.query_to_ref, d.q2r_runs = None, runs.ctypes.data
            else:
                q2r = np.ascontiguousarray(r.queryToRef, dtype=np.int32)
                keep.append(q2r)
                d.query_to_ref = q2r.ctypes.data
            keep.append((r.basecall, r.referenceSeqMappedTo))
        return arr, keep

    def submit(self, reads) -> Batch:
        arr, keep = self._descs(reads)
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit(self.h, C.addressof(arr), len(reads), C.byref(h)), "dnb_submit")
        return Batch(self, h, len(reads), (arr, keep))

    def upload(self, reads) -> Batch:
        arr, keep = self._descs(reads)
        h = C.c_void_p()
        _lib.check(self.L.dnb_batch_upload(self.h, C.addressof(arr), len(reads), C.byref(h)), "dnb_batch_upload")
        return Batch(self, h, len(reads), None)   # inputs are resident in HBM; host copies may be dropped

    def submit_chain(self, reads, extra, window: int = 50, want_records: bool = False) -> Batch:
        """dnb_submit_chain: normaliseEvents + eventalign + DNN input tensors in one pipelined call; read the results with
        Batch.results() and Batch.feature_results()."""
        descs, keep = self._descs(reads)
        ex, keep2 = _extras(extra)
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit_chain(self.h, C.addressof(descs), C.addressof(ex), len(reads), window,
                                           int(want_records), C.byref(h)), "dnb_submit_chain")
        return Batch(self, h, len(reads), (keep, keep2))

    def submit_llr(self, reads, extra, window: int = 12) -> Batch:
        """dnb_submit_llr: normaliseEvents + llAcrossRead (the --HMM read loop body, detect.cpp:876-885) in one pipelined
        call; read the results with Batch.results() and Batch.analogue_results()."""
        descs, keep = self._descs(reads)
        ex, keep2 = _extras(extra)
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit_llr(self.h, C.addressof(descs), C.addressof(ex), len(reads), window, C.byref(h)),
                   "dnb_submit_llr")
        return Batch(self, h, len(reads), (keep, keep2))

    def submit_llr_descs(self, descs: np.ndarray, extras: np.ndarray, window: int = 12) -> Batch:
        assert descs.size == extras.size
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit_llr(self.h, descs.ctypes.data, extras.ctypes.data, descs.size, window, C.byref(h)),
                   "dnb_submit_llr")
        return Batch(self, h, descs.size, (descs, extras))

    # descriptor arrays built with numpy (dtype _lib.READ_DESC_DTYPE): no per-read Python objects
    def submit_descs(self, descs: np.ndarray) -> Batch:
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit(self.h, descs.ctypes.data, descs.size, C.byref(h)), "dnb_submit")
        return Batch(self, h, descs.size, descs)

    def submit_chain_descs(self, descs: np.ndarray, extras: np.ndarray, window: int = 50, want_records: bool = False) -> Batch:
        """dnb_submit_chain on descriptor arrays (dtypes _lib.READ_DESC_DTYPE / _lib.READ_EXTRA_DTYPE)."""
        assert descs.size == extras.size
        h = C.c_void_p()
        _lib.check(self.L.dnb_submit_chain(self.h, descs.ctypes.data, extras.ctypes.data, descs.size, window,
                                           int(want_records), C.byref(h)), "dnb_submit_chain")
        return Batch(self, h, descs.size, (descs, extras))

    def upload_descs(self, descs: np.ndarray) -> Batch:
        h = C.c_void_p()
        _lib.check(self.L.dnb_batch_upload(self.h, descs.ctypes.data, descs.size, C.byref(h)), "dnb_batch_upload")
        return Batch(self, h, descs.size, None)

    def normaliseEvents(self, reads) -> list[Normalised]:
        """Batched normaliseEvents(r, false): src/event_handling.cpp:544."""
        b = self.submit(reads)
        try:
            b.wait()
            return b.results()
        finally:
            b.release()

    # -- detect_events ------------------------------------------------------------------------------
    def detect_events(self, raw: np.ndarray):
        """event_table detect_events(raw, n, event_detection_defaults): src/scrappie/event_detection.c:268."""
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        cap = raw.size + 2
        ev = (EventT * cap)()
        n = C.c_size_t(0)
        _lib.check(self.L.dnb_detect_events(self.h, raw.ctypes.data, raw.size, ev, cap, C.byref(n)), "dnb_detect_events")
        a = np.ctypeslib.as_array(ev)[: n.value]
        return (a["start"].astype(np.uint64), a["length"].astype(np.float32), a["mean"].astype(np.float32),
                a["stdv"].astype(np.float32))

    # -- Theil-Sen refinement alone ------------------------------------------------------------------
    def theil_sen_batch(self, signals_list, ranks_list, rough_shift, rough_scale):
        """estimateScaling_theilSen (src/event_handling.cpp:24-110) for every read's cleaned (signal, k-mer rank) vectors;
        returns (shift, scale) arrays."""
        n = len(signals_list)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in signals_list])
        sig = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in signals_list]) if n else np.zeros(0))
        rk = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint32) for x in ranks_list]) if n else np.zeros(0, np.uint32))
        rs = np.ascontiguousarray(np.broadcast_to(rough_shift, (n,)), dtype=np.float64)
        rc = np.ascontiguousarray(np.broadcast_to(rough_scale, (n,)), dtype=np.float64)
        shift, scale = np.zeros(n), np.zeros(n)
        _lib.check(self.L.dnb_theil_sen_batch(self.h, sig.ctypes.data, rk.ctypes.data, off.ctypes.data, n, rs.ctypes.data,
                                              rc.ctypes.data, shift.ctypes.data, scale.ctypes.data), "dnb_theil_sen_batch")
        return shift, scale

    # -- analogue likelihood ------------------------------------------------------------------------
    def sequence_probability_batch(self, obs_list, snippets, shift, scale, events_per_base, window: int = 12):
        """Both passes of sequenceProbability for every site (src/detect.cpp:235-378, 546-547)."""
        n = len(obs_list)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(o) for o in obs_list])
        obs = np.ascontiguousarray(np.concatenate([np.asarray(o, dtype=np.float64) for o in obs_list])
                                   if n else np.zeros(0), dtype=np.float64)
        seq = b"".join(snippets)
        assert len(seq) == n * (2 * window + 9)
        shift = np.ascontiguousarray(np.broadcast_to(shift, (n,)), dtype=np.float64)
        scale = np.ascontiguousarray(np.broadcast_to(scale, (n,)), dtype=np.float64)
        epb = np.ascontiguousarray(np.broadcast_to(events_per_base, (n,)), dtype=np.float64)
        oa = np.zeros(n)
        ot = np.zeros(n)
        _lib.check(self.L.dnb_sequence_probability_batch(self.h, obs.ctypes.data, off.ctypes.data, seq, shift.ctypes.data,
                                                         scale.ctypes.data, epb.ctypes.data, n, window, oa.ctypes.data,
                                                         ot.ctypes.data), "dnb_sequence_probability_batch")
        return oa, ot


    # -- eventalign (src/alignment.cpp:547-744) -------------------------------------------------------------
    def eventalign(self, reads, window: int = 50):
        """Batched eventalign.  `reads`: dicts with refseq (bytes), ref_to_query (int32[ref_len]), eventAlignment
        (uint32[n,2]), event_mean (float32[n_events]), shift, scale, events_per_base -- the fields eventalign reads
        from DNAscent::read after normaliseEvents.  Returns per read a dict of record arrays (event, ref_pos, label
        1 = M / 2 = I, indel) and a status; formatting the text / addSignal is the host's job (shim)."""
        n = len(reads)
        descs = (_lib.EventalignDesc * max(n, 1))()
        keep = []
        rec_off = np.zeros(n + 1, dtype=np.uint64)
        for i, r in enumerate(reads):
            ref = np.frombuffer(r["refseq"], dtype=np.uint8)
            r2q = np.ascontiguousarray(r["ref_to_query"], dtype=np.int32)
            al = np.ascontiguousarray(r["eventAlignment"], dtype=np.uint32).reshape(-1, 2)
            evm = np.ascontiguousarray(r["event_mean"], dtype=np.float32)
            assert r2q.size >= ref.size
            keep += [ref, r2q, al, evm]
            d = descs[i]
            d.ref, d.ref_len, d.ref_to_query = ref.ctypes.data, ref.size, r2q.ctypes.data
            d.align_pairs, d.n_align = al.ctypes.data, al.shape[0]
            d.event_mean, d.n_events = evm.ctypes.data, evm.size
            d.shift, d.scale, d.events_per_base = float(r["shift"]), float(r["scale"]), float(r["events_per_base"])
            rec_off[i + 1] = rec_off[i] + al.shape[0] + 64
        recs = np.zeros(int(rec_off[n]), dtype=_lib.EVENTALIGN_REC_DTYPE)
        n_rec = np.zeros(max(n, 1), dtype=np.uint32)
        status = np.zeros(max(n, 1), dtype=np.int32)
        _lib.check(self.L.dnb_eventalign_batch(self.h, C.addressof(descs), n, window, recs.ctypes.data,
                                               rec_off.ctypes.data, n_rec.ctypes.data, status.ctypes.data),
                   "dnb_eventalign_batch")
        out = []
        for i in range(n):
            rr = recs[int(rec_off[i]):int(rec_off[i]) + int(n_rec[i])]
            out.append(dict(event=rr["event"].copy(), ref_pos=rr["ref_pos"].copy(), label=rr["label"].astype(np.uint8),
                            indel=rr["indel_score"].copy(), status=int(status[i])))
        return out

    def eventalign_last_kernel_ms(self) -> float:
        return float(self.L.dnb_eventalign_last_kernel_ms())

    # -- eventalign + DNN input tensors (src/reads.h:288-452, consumer src/detect.cpp:586-649) -----------------
    def eventalign_features(self, reads, window: int = 50, want_records: bool = True):
        """eventalign and, in the same device pass, the tensors runCNN feeds TensorFlow.  `reads`: the dicts
        eventalign() takes plus raw (float32 pA) or raw_dac (int16) with dac_offset / dac_scale, event_start
        (uint32[n_events+1]), is_reverse, ref_start, ref_end and optionally called (sorted uint32 coordinates).
        Returns per read a dict: status, signal [P,20], core, residual, coords, ref_index, query_index, quality
        (+ the record arrays when want_records)."""
        n = len(reads)
        descs = (_lib.EventalignDesc * max(n, 1))()
        feats = (_lib.FeatureDesc * max(n, 1))()
        keep = []
        rec_off = np.zeros(n + 1, dtype=np.uint64)
        pos_off = np.zeros(n + 1, dtype=np.uint64)
        for i, r in enumerate(reads):
            ref = np.frombuffer(r["refseq"], dtype=np.uint8)
            r2q = np.ascontiguousarray(r["ref_to_query"], dtype=np.int32)
            al = np.ascontiguousarray(r["eventAlignment"], dtype=np.uint32).reshape(-1, 2)
            evm = np.ascontiguousarray(r["event_mean"], dtype=np.float32)
            es = np.ascontiguousarray(r["event_start"], dtype=np.uint32)
            called = np.ascontiguousarray(r.get("called", []), dtype=np.uint32)
            assert r2q.size >= ref.size and es.size == evm.size + 1
            keep += [ref, r2q, al, evm, es, called]
            d, f = descs[i], feats[i]
            d.ref, d.ref_len, d.ref_to_query = ref.ctypes.data, ref.size, r2q.ctypes.data
            d.align_pairs, d.n_align = al.ctypes.data, al.shape[0]
            d.event_mean, d.n_events = evm.ctypes.data, evm.size
            d.shift, d.scale, d.events_per_base = float(r["shift"]), float(r["scale"]), float(r["events_per_base"])
            if r.get("raw_dac") is not None:
                raw = np.ascontiguousarray(r["raw_dac"], dtype=np.int16)
                f.raw_dac, f.dac_offset, f.dac_scale = raw.ctypes.data, float(r["dac_offset"]), float(r["dac_scale"])
            else:
                raw = np.ascontiguousarray(r["raw"], dtype=np.float32)
                f.raw_pA = raw.ctypes.data
            keep.append(raw)
            f.n_samples, f.event_start = raw.size, es.ctypes.data
            f.is_reverse, f.ref_start, f.ref_end = int(bool(r["is_reverse"])), int(r["ref_start"]), int(r["ref_end"])
            f.called, f.n_called = (called.ctypes.data if called.size else None), called.size
            rec_off[i + 1] = rec_off[i] + al.shape[0] + 64
            pos_off[i + 1] = pos_off[i] + int(r.get("row_capacity", max(ref.size, 8) - 8 + 1))
        recs = np.zeros(int(rec_off[n]) if want_records else 0, dtype=_lib.EVENTALIGN_REC_DTYPE)
        n_rec = np.zeros(max(n, 1), dtype=np.uint32)
        n_pos = np.zeros(max(n, 1), dtype=np.uint32)
        status = np.zeros(max(n, 1), dtype=np.int32)
        rows = int(pos_off[n])
        T = dict(signal=np.zeros((rows, _lib.RAWDEPTH), dtype=np.float32), core=np.zeros(rows, dtype=np.float32),
                 residual=np.zeros(rows, dtype=np.float32), coords=np.zeros(rows, dtype=np.uint32),
                 ref_index=np.zeros(rows, dtype=np.uint32), query_index=np.zeros(rows, dtype=np.uint32),
                 quality=np.zeros(rows, dtype=np.int32))
        tens = _lib.FeatureTensors(**{k: v.ctypes.data for k, v in T.items()})
        _lib.check(self.L.dnb_eventalign_features_batch(
            self.h, C.addressof(descs), C.addressof(feats), n, window, recs.ctypes.data if want_records else None,
            rec_off.ctypes.data, n_rec.ctypes.data, status.ctypes.data, C.addressof(tens), pos_off.ctypes.data,
            n_pos.ctypes.data), "dnb_eventalign_features_batch")
        out = []
        for i in range(n):
            lo, hi = int(pos_off[i]), int(pos_off[i]) + int(n_pos[i])
            o = {k: v[lo:hi].copy() for k, v in T.items()}
            o["status"] = int(status[i])
            if want_records:
                rr = recs[int(rec_off[i]):int(rec_off[i]) + int(n_rec[i])]
                o.update(event=rr["event"].copy(), ref_pos=rr["ref_pos"].copy(), label=rr["label"].astype(np.uint8),
                         indel=rr["indel_score"].copy())
            out.append(o)
        return out

    def features_last_kernel_ms(self) -> float:
        return float(self.L.dnb_features_last_kernel_ms())


def host_stats(reset: bool = False):
    """dnb_host_stats: ({phase name: thread-seconds}, {counter: n}) of the batch pipeline's host side since the last reset."""
    L = _lib.lib()
    sec = (C.c_double * 17)()
    cnt = (C.c_uint64 * 4)()
    _lib.check(L.dnb_host_stats(int(reset), C.byref(sec), C.byref(cnt)), "dnb_host_stats")
    names = [L.dnb_host_phase_name(i).decode() for i in range(17)]
    return (dict(zip(names, (float(x) for x in sec))),
            dict(zip(("batches", "direct_dma_batches", "cudaMalloc_calls", "cudaMallocHost_calls"), (int(x) for x in cnt))))


def host_register(arr: np.ndarray):
    """Page-lock a numpy buffer in place (dnb_host_register): dnb_submit then DMAs reads out of it directly."""
    _lib.check(_lib.lib().dnb_host_register(arr.ctypes.data, arr.nbytes), "dnb_host_register")


def host_unregister(arr: np.ndarray):
    _lib.check(_lib.lib().dnb_host_unregister(arr.ctypes.data), "dnb_host_unregister")


def dorado_slice(n_total: int, signal_length: int = 0, signal_trim: int = 0, signal_start_coord: int = 0,
                 is_split: bool = False) -> slice:
    """The part of a POD5 record's signal that pod5_getSignal keeps (src/pod5.cpp:76-93), as a Python slice to apply
    to the int16 DAC array before it goes into Read(dac=...).  Raises where the reference's erase is undefined."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    _lib.check(_lib.lib().dnb_dorado_slice(n_total, signal_length, signal_trim, signal_start_coord, int(is_split),
                                           C.byref(a), C.byref(b)), "dnb_dorado_slice")
    return slice(a.value, a.value + b.value)


# ---- probability.h drop-ins (scalar, host) ------------------------------------------------------------
def eexp(x: float) -> float:
    return _lib.lib().dnb_eexp(x)


def eln(x: float) -> float:
    o = C.c_double(0.0)
    rc = _lib.lib().dnb_eln(x, C.byref(o))
    if rc == ERR_NEGATIVE_LOG:
        raise NegativeLog()
    _lib.check(rc, "dnb_eln")
    return o.value


def lnSum(a: float, b: float) -> float:
    return _lib.lib().dnb_lnSum(a, b)


def lnProd(a: float, b: float) -> float:
    return _lib.lib().dnb_lnProd(a, b)


def lnGreaterThan(a: float, b: float) -> bool:
    return bool(_lib.lib().dnb_lnGreaterThan(a, b))


def uniformPDF(lb: float, ub: float, x: float) -> float:
    return _lib.lib().dnb_uniformPDF(lb, ub, x)


def normalPDF(mu: float, sigma: float, x: float) -> float:
    return _lib.lib().dnb_normalPDF(mu, sigma, x)


def cauchyPDF(loc: float, scale: float, x: float) -> float:
    return _lib.lib().dnb_cauchyPDF(loc, scale, x)


# ---- llAcrossRead helper: the event gathering around every T of the reference (host side) -----------------------
def gather_sites(refseq: bytes, ref_to_query, is_reverse: bool, event_alignment, event_mean, window: int = 12):
    """Per T site of `referenceSeqMappedTo`, the events sequenceProbability is called on: a transcription of the
    gathering loop of llAcrossRead (src/detect.cpp:381-390, 399-510).  Returns [(posOnRef, events, snippet)].
    The forward passes themselves run on the device (Context.sequence_probability_batch)."""
    k = 9
    n = len(refseq)
    out = []
    if n < 4 * window + 1 or len(event_alignment) == 0:
        return out
    r2q = np.asarray(ref_to_query)
    al_e = np.asarray(event_alignment)[:, 0]
    al_k = np.asarray(event_alignment)[:, 1]
    pois = [i for i in range(2 * window, n - 2 * window) if refseq[i:i + 1] == b"T"]
    read_head = 0
    if is_reverse:
        read_head = len(al_e) - 1
        pois.reverse()
    for p in pois:
        sn = refseq[p - window:p + window + k]
        if len(sn) != 2 * window + k or any(c not in b"ATGC" for c in sn):
            continue
        lo, hi = int(r2q[p - window]) & 0xFFFFFFFF, int(r2q[p + window]) & 0xFFFFFFFF
        ev, first = [], True
        if is_reverse:
            j = read_head
            while j >= 0:
                if lo <= al_k[j] < hi:
                    if first:
                        read_head, first = j, False
                    m = float(event_mean[al_e[j]])
                    if 0.0 < m < 250.0:
                        ev.append(m)
                if al_k[j] < lo:
                    ev.reverse()
                    break
                j -= 1
        else:
            for j in range(read_head, len(al_e)):
                if lo <= al_k[j] < hi:
                    if first:
                        read_head, first = j, False
                    m = float(event_mean[al_e[j]])
                    if 0.0 < m < 250.0:
                        ev.append(m)
                if al_k[j] >= hi:
                    break
        if len(ev) < 2 * window - k:
            continue
        out.append((p, np.array(ev), sn))
    return out
