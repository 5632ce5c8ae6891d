"""TEST INFRASTRUCTURE: ctypes binding of the plain-C oracle port (oracle/dnb_oracle.c ->
oracle/_build/libdnb_oracle.so).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libdnb_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "dnb_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "dnb_oracle.h"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB_PATH


class _Result(C.Structure):
    _fields_ = [
        ("et_n", C.c_size_t), ("et_start", C.POINTER(C.c_uint64)), ("et_length", C.POINTER(C.c_float)),
        ("et_mean", C.POINTER(C.c_float)), ("et_stdv", C.POINTER(C.c_float)),
        ("n_events", C.c_size_t), ("ev_mean", C.POINTER(C.c_double)), ("ev_start", C.POINTER(C.c_uint32)),
        ("n_kmers", C.c_size_t), ("n_kmers_ref", C.c_size_t),
        ("rank_query", C.POINTER(C.c_uint32)), ("rank_ref", C.POINTER(C.c_uint32)),
        ("rough_shift", C.c_double), ("rough_scale", C.c_double), ("shift", C.c_double), ("scale", C.c_double),
        ("events_per_base", C.c_double),
        ("n_bands", C.c_size_t), ("band_move", C.POINTER(C.c_uint8)), ("trace", C.POINTER(C.c_uint8)),
        ("last_col", C.POINTER(C.c_float)),
        ("lp_skip", C.c_double), ("lp_stay", C.c_double), ("lp_step", C.c_double), ("lp_trim", C.c_double),
        ("fills", C.c_int64),
        ("n_align", C.c_size_t), ("align_event", C.POINTER(C.c_uint32)), ("align_kmer", C.POINTER(C.c_uint32)),
        ("n_cleaned", C.c_size_t), ("cleaned_signal", C.POINTER(C.c_double)), ("cleaned_rank", C.POINTER(C.c_uint32)),
        ("avg_log_emission", C.c_double), ("spanned", C.c_int), ("max_gap", C.c_int), ("end_score", C.c_float),
        ("status", C.c_int),
    ]


class _DetParam(C.Structure):
    _fields_ = [("w1", C.c_uint32), ("w2", C.c_uint32), ("thr1", C.c_float), ("thr2", C.c_float),
                ("peak_height", C.c_float)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


OK, QC_FAIL, SCALE_FAIL, UNDEFINED = 0, 1, 2, 3


class Port:
    def __init__(self):
        L = self.L = C.CDLL(build())
        sz, vp, d = C.c_size_t, C.c_void_p, C.c_double
        L.dnbo_detect_events.restype = sz
        L.dnbo_detect_events.argtypes = [vp, sz, _DetParam, vp, vp, vp, vp, sz]
        L.dnbo_tstat.argtypes = [vp, sz, C.c_uint32, vp]
        L.dnbo_kmer2index.restype = C.c_uint32
        L.dnbo_kmer2index.argtypes = [C.c_char_p, C.c_uint]
        L.dnbo_kmer_ranks.argtypes = [C.c_char_p, sz, vp]
        L.dnbo_quantile_scaling.argtypes = [vp, sz, vp, sz, vp, vp, vp]
        L.dnbo_log_probability_match.restype = C.c_float
        L.dnbo_log_probability_match.argtypes = [d, d, d, d, d]
        L.dnbo_theil_sen.argtypes = [vp, vp, sz, vp, d, d, vp, vp]
        L.dnbo_normalise.argtypes = [vp, sz, C.c_char_p, sz, C.c_char_p, sz, vp, vp, vp, C.c_int, C.POINTER(_Result)]
        L.dnbo_result_free.argtypes = [C.POINTER(_Result)]
        L.dnbo_bench_normalise.restype = d
        L.dnbo_bench_normalise.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, vp, C.c_int, vp]
        L.dnbo_eexp.restype = d
        L.dnbo_eexp.argtypes = [d]
        L.dnbo_eln.argtypes = [d, vp]
        for f in ("dnbo_lnSum", "dnbo_lnProd"):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d, d]
        L.dnbo_lnGreaterThan.argtypes = [d, d]
        for f in ("dnbo_uniformPDF", "dnbo_normalPDF", "dnbo_cauchyPDF"):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d, d, d]
        L.dnbo_sequence_probability.restype = d
        L.dnbo_sequence_probability.argtypes = [vp, sz, C.c_char_p, sz, sz, C.c_int, d, d, d, sz, sz, vp, vp, vp, vp]
        L.dnbo_builtin_viterbi.restype = sz
        L.dnbo_builtin_viterbi.argtypes = [vp, sz, C.c_char_p, sz, d, d, d, vp, vp, vp, vp, vp, sz]
        L.dnbo_eventalign.restype = sz
        L.dnbo_eventalign.argtypes = [C.c_char_p, sz, vp, vp, vp, sz, vp, d, d, d, C.c_uint, vp, vp, vp, vp, vp, sz]
        L.dnbo_dnn_features.restype = sz
        L.dnbo_dnn_features.argtypes = [C.c_char_p, sz, vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp, vp, sz, vp, vp,
                                        d, d, vp, sz, vp, vp, vp, vp, vp, vp, vp, sz]
        L.dnbo_ll_across_read.restype = sz
        L.dnbo_ll_across_read.argtypes = [C.c_char_p, sz, vp, C.c_int, vp, vp, sz, vp, d, d, d, C.c_uint, vp, vp, vp,
                                          vp, vp, vp, sz]

    def dnn_features(self, ref: bytes, r2q, is_reverse, ref_start, ref_end, rec, raw, event_start, shift, scale,
                     called=None):
        """DNN input tensors from eventalign records + raw signal -- reads.h:288-372, alignment.cpp:706-725"""
        r2q = np.ascontiguousarray(r2q, dtype=np.int32)
        ev = np.ascontiguousarray(rec["event"], dtype=np.uint32)
        rp = np.ascontiguousarray(rec["ref_pos"], dtype=np.uint32)
        lb = np.ascontiguousarray(rec["label"], dtype=np.uint8)
        ind = np.ascontiguousarray(rec["indel"], dtype=np.int32)
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        es = np.ascontiguousarray(event_start, dtype=np.uint32)
        called = np.ascontiguousarray(called if called is not None else [], dtype=np.uint32)
        cap = len(ref) + 1
        sig = np.zeros((cap, 20), dtype=np.float32)
        core = np.zeros(cap, dtype=np.float32)
        resid = np.zeros(cap, dtype=np.float32)
        coords = np.zeros(cap, dtype=np.uint32)
        ri = np.zeros(cap, dtype=np.uint32)
        qi = np.zeros(cap, dtype=np.uint32)
        qual = np.zeros(cap, dtype=np.int32)
        P = self.L.dnbo_dnn_features(ref, len(ref), _p(r2q), int(is_reverse), int(ref_start), int(ref_end), _p(ev),
                                     _p(rp), _p(lb), _p(ind), ev.size, _p(raw), _p(es), shift, scale, _p(called),
                                     called.size, _p(sig), _p(core), _p(resid), _p(coords), _p(ri), _p(qi), _p(qual),
                                     cap)
        assert P <= cap
        return dict(signal=sig[:P].copy(), core=core[:P].copy(), residual=resid[:P].copy(), coords=coords[:P].copy(),
                    ref_index=ri[:P].copy(), query_index=qi[:P].copy(), quality=qual[:P].copy())

    def builtin_viterbi(self, obs, seq: bytes, shift, scale, epb, model_mean, model_stdv=None):
        """(score, state index[], state type[] 0=D 1=M 2=I) -- alignment.cpp:193-516"""
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        cap = 4 * (obs.size + len(seq)) + 16
        idx = np.zeros(cap, dtype=np.int32)
        typ = np.zeros(cap, dtype=np.uint8)
        score = C.c_double(0.0)
        n = self.L.dnbo_builtin_viterbi(_p(obs), obs.size, seq, len(seq), shift, scale, epb, _p(model_mean),
                                        _p(model_stdv) if model_stdv is not None else None, C.byref(score), _p(idx),
                                        _p(typ), cap)
        return score.value, idx[:n].copy(), typ[:n].copy()

    def eventalign(self, ref: bytes, r2q, align_event, align_kmer, ev_mean, shift, scale, epb, model_mean, window=50):
        """records (event, ref_pos, label 1=M 2=I, indelScore) -- alignment.cpp:547-744"""
        r2q = np.ascontiguousarray(r2q, dtype=np.int32)
        ae = np.ascontiguousarray(align_event, dtype=np.uint32)
        ak = np.ascontiguousarray(align_kmer, dtype=np.uint32)
        ev_mean = np.ascontiguousarray(ev_mean, dtype=np.float64)
        cap = ae.size + 16
        ev = np.zeros(cap, dtype=np.uint32)
        rp = np.zeros(cap, dtype=np.uint32)
        lb = np.zeros(cap, dtype=np.uint8)
        ind = np.zeros(cap, dtype=np.int32)
        n = self.L.dnbo_eventalign(ref, len(ref), _p(r2q), _p(ae), _p(ak), ae.size, _p(ev_mean), shift, scale, epb,
                                   window, _p(model_mean), _p(ev), _p(rp), _p(lb), _p(ind), cap)
        assert n <= cap
        return dict(event=ev[:n].copy(), ref_pos=rp[:n].copy(), label=lb[:n].copy(), indel=ind[:n].copy())

    def detect_events(self, raw, w1=3, w2=6, thr1=1.4, thr2=9.0, peak_height=0.2):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        cap = raw.size + 2
        start = np.zeros(cap, dtype=np.uint64)
        length = np.zeros(cap, dtype=np.float32)
        mean = np.zeros(cap, dtype=np.float32)
        stdv = np.zeros(cap, dtype=np.float32)
        n = self.L.dnbo_detect_events(_p(raw), raw.size, _DetParam(w1, w2, thr1, thr2, peak_height), _p(start),
                                      _p(length), _p(mean), _p(stdv), cap)
        return start[:n], length[:n], mean[:n], stdv[:n]

    def tstat(self, raw, w):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        out = np.zeros(raw.size, dtype=np.float32)
        self.L.dnbo_tstat(_p(raw), raw.size, w, _p(out))
        return out

    def kmer_ranks(self, seq: bytes):
        out = np.zeros(max(len(seq) - 8, 0), dtype=np.uint32)
        self.L.dnbo_kmer_ranks(seq, len(seq), _p(out))
        return out

    def quantile_scaling(self, ev_mean, rank_ref, model_mean):
        ev_mean = np.ascontiguousarray(ev_mean, dtype=np.float64)
        rank_ref = np.ascontiguousarray(rank_ref, dtype=np.uint32)
        o = np.zeros(2)
        rc = self.L.dnbo_quantile_scaling(_p(ev_mean), ev_mean.size, _p(rank_ref), rank_ref.size, _p(model_mean),
                                          _p(o[0:1]), _p(o[1:2]))
        return rc, float(o[0]), float(o[1])

    def theil_sen(self, sig, ranks, model_mean, shift, scale):
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
        o = np.zeros(2)
        self.L.dnbo_theil_sen(_p(sig), _p(ranks), sig.size, _p(model_mean), shift, scale, _p(o[0:1]), _p(o[1:2]))
        return float(o[0]), float(o[1])

    def normalise(self, raw, query: bytes, ref: bytes, q2r, model_mean, model_stdv=None, keep_bands=False) -> dict:
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        q2r = np.ascontiguousarray(q2r, dtype=np.int32)
        model_mean = np.ascontiguousarray(model_mean, dtype=np.float64)
        res = _Result()
        sd = _p(np.ascontiguousarray(model_stdv, dtype=np.float64)) if model_stdv is not None else None
        rc = self.L.dnbo_normalise(_p(raw), raw.size, query, len(query), ref, len(ref), _p(q2r), _p(model_mean), sd,
                                   int(keep_bands), C.byref(res))
        E, na, nc = res.n_events, res.n_align, res.n_cleaned
        out = dict(
            status=rc, et_n=res.et_n, et_start=_np(res.et_start, res.et_n, np.uint64),
            et_length=_np(res.et_length, res.et_n, np.float32), et_mean=_np(res.et_mean, res.et_n, np.float32),
            et_stdv=_np(res.et_stdv, res.et_n, np.float32),
            event_mean=_np(res.ev_mean, E, np.float64), event_start=_np(res.ev_start, E + 1 if res.ev_start else 0, np.uint32),
            rank_query=_np(res.rank_query, res.n_kmers, np.uint32), rank_ref=_np(res.rank_ref, res.n_kmers_ref, np.uint32),
            rough_shift=res.rough_shift, rough_scale=res.rough_scale, shift=res.shift, scale=res.scale,
            events_per_base=res.events_per_base, n_bands=res.n_bands, fills=res.fills,
            lp=(res.lp_skip, res.lp_stay, res.lp_step, res.lp_trim),
            align_event=_np(res.align_event, na, np.uint32), align_kmer=_np(res.align_kmer, na, np.uint32),
            cleaned_signal=_np(res.cleaned_signal, nc, np.float64), cleaned_rank=_np(res.cleaned_rank, nc, np.uint32),
            avg_log_emission=res.avg_log_emission, spanned=bool(res.spanned), max_gap=res.max_gap,
            end_score=res.end_score)
        if keep_bands and res.trace:
            out["band_move"] = _np(res.band_move, res.n_bands, np.uint8)
            out["trace"] = _np(res.trace, res.n_bands * 100, np.uint8).reshape(res.n_bands, 100)
            out["last_col"] = _np(res.last_col, E, np.float32)
        self.L.dnbo_result_free(C.byref(res))
        return out

    def bench_normalise(self, reads, model_mean, threads: int):
        """reads: objects with .raw (float32), .basecall, .refseq, .query_to_ref"""
        n = len(reads)
        raws = [np.ascontiguousarray(r.raw, dtype=np.float32) for r in reads]
        q2rs = [np.ascontiguousarray(r.query_to_ref, dtype=np.int32) for r in reads]
        qs = [C.c_char_p(r.basecall) for r in reads]
        rs = [C.c_char_p(r.refseq) for r in reads]
        vp, sz = C.c_void_p, C.c_size_t
        a_raw = (vp * n)(*[x.ctypes.data for x in raws])
        a_nraw = (sz * n)(*[x.size for x in raws])
        a_q = (C.c_char_p * n)(*qs)
        a_ql = (sz * n)(*[len(r.basecall) for r in reads])
        a_r = (C.c_char_p * n)(*rs)
        a_rl = (sz * n)(*[len(r.refseq) for r in reads])
        a_q2r = (vp * n)(*[x.ctypes.data for x in q2rs])
        failed = C.c_int(0)
        mm = np.ascontiguousarray(model_mean, dtype=np.float64)
        t = self.L.dnbo_bench_normalise(a_raw, a_nraw, a_q, a_ql, a_r, a_rl, a_q2r, n, _p(mm), threads,
                                        C.byref(failed))
        return t, failed.value

    def eln(self, x):
        o = np.zeros(1)
        if self.L.dnbo_eln(x, _p(o)):
            raise ValueError("NegativeLog")
        return float(o[0])

    def sequence_probability(self, obs, seq: bytes, window, use_analogue, shift, scale, epb, a_start, a_end,
                             unl_mean, unl_stdv, ana_mean, ana_stdv) -> float:
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        return self.L.dnbo_sequence_probability(_p(obs), obs.size, seq, len(seq), window, int(use_analogue), shift,
                                                scale, epb, a_start, a_end, _p(unl_mean), _p(unl_stdv), _p(ana_mean),
                                                _p(ana_stdv))

    def ll_across_read(self, ref: bytes, r2q, is_reverse, align_event, align_kmer, ev_mean, shift, scale, epb, window,
                       unl_mean, unl_stdv, ana_mean, ana_stdv):
        r2q = np.ascontiguousarray(r2q, dtype=np.int32)
        ae = np.ascontiguousarray(align_event, dtype=np.uint32)
        ak = np.ascontiguousarray(align_kmer, dtype=np.uint32)
        ev_mean = np.ascontiguousarray(ev_mean, dtype=np.float64)
        cap = len(ref) + 1
        pos = np.zeros(cap, dtype=np.uint32)
        llr = np.zeros(cap)
        n = self.L.dnbo_ll_across_read(ref, len(ref), _p(r2q), int(is_reverse), _p(ae), _p(ak), ae.size, _p(ev_mean),
                                       shift, scale, epb, window, _p(unl_mean), _p(unl_stdv), _p(ana_mean),
                                       _p(ana_stdv), _p(pos), _p(llr), cap)
        return pos[:n].copy(), llr[:n].copy()
