"""ctypes loader for the in-tree CUDA library.  There is no fallback: if the shared object is missing the import
of anything that computes fails loudly (the oracle under oracle/ is test infrastructure and is never used here)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdnascent_b200.so")

_lib = None


class DnbError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: error {code}" + (f" ({detail})" if detail else ""))


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("window_length1", C.c_uint32), ("window_length2", C.c_uint32),
        ("threshold1", C.c_float), ("threshold2", C.c_float), ("peak_height", C.c_float),
        ("min_average_log_emission", C.c_double), ("max_gap_threshold", C.c_int), ("bandwidth", C.c_int),
        ("use_fit_pore_model", C.c_int), ("event_capacity_per_sample", C.c_float), ("keep_debug", C.c_int),
        ("workspace_bytes", C.c_size_t),
        ("result_format", C.c_int), ("n_devices", C.c_int), ("devices", C.c_int * 16),
    ]


class ReadDesc(C.Structure):
    _fields_ = [
        ("raw_pA", C.c_void_p), ("raw_dac", C.c_void_p), ("dac_offset", C.c_float), ("dac_scale", C.c_float),
        ("n_samples", C.c_uint64),
        ("query", C.c_char_p), ("query_len", C.c_uint32),
        ("ref", C.c_char_p), ("ref_len", C.c_uint32),
        ("query_to_ref", C.c_void_p),
        ("q2r_runs", C.c_void_p), ("n_q2r_runs", C.c_uint32),
    ]


class Q2RRun(C.Structure):
    _fields_ = [("q_start", C.c_uint32), ("len", C.c_uint32), ("r_start", C.c_int32), ("stride", C.c_int32)]


# numpy mirror of dnb_read_desc (C layout, natural alignment) for building descriptor arrays without a Python loop
import numpy as _np
READ_DESC_DTYPE = _np.dtype([
    ("raw_pA", _np.uint64), ("raw_dac", _np.uint64), ("dac_offset", _np.float32), ("dac_scale", _np.float32),
    ("n_samples", _np.uint64), ("query", _np.uint64), ("query_len", _np.uint32), ("ref", _np.uint64),
    ("ref_len", _np.uint32), ("query_to_ref", _np.uint64), ("q2r_runs", _np.uint64), ("n_q2r_runs", _np.uint32)], align=True)
Q2R_RUN_DTYPE = _np.dtype([("q_start", _np.uint32), ("len", _np.uint32), ("r_start", _np.int32), ("stride", _np.int32)])
assert Q2R_RUN_DTYPE.itemsize == C.sizeof(Q2RRun)
assert READ_DESC_DTYPE.itemsize == C.sizeof(ReadDesc), (READ_DESC_DTYPE.itemsize, C.sizeof(ReadDesc))


class ReadResult(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("et_n", C.c_uint32), ("n_events", C.c_uint32),
        ("event_start", C.POINTER(C.c_uint32)), ("event_mean", C.POINTER(C.c_float)),
        ("n_align", C.c_uint32), ("align_pairs", C.POINTER(C.c_uint32)),
        ("shift", C.c_double), ("scale", C.c_double), ("events_per_base", C.c_double),
        ("rough_shift", C.c_double), ("rough_scale", C.c_double),
        ("avg_log_emission", C.c_double), ("spanned", C.c_int), ("max_gap", C.c_int),
        ("n_cleaned", C.c_uint32), ("cleaned_signal", C.POINTER(C.c_double)), ("cleaned_rank", C.POINTER(C.c_uint32)),
        ("event_first", C.c_uint32), ("event_len8", C.POINTER(C.c_uint8)), ("event_len_escape", C.POINTER(C.c_uint32)),
        ("n_event_len_escape", C.c_uint32), ("align_first", C.c_uint32 * 2), ("align_steps", C.POINTER(C.c_uint8)),
    ]


class EventalignDesc(C.Structure):
    _fields_ = [
        ("ref", C.c_void_p), ("ref_len", C.c_uint32), ("ref_to_query", C.c_void_p),
        ("align_pairs", C.c_void_p), ("n_align", C.c_uint32),
        ("event_mean", C.c_void_p), ("n_events", C.c_uint32),
        ("shift", C.c_double), ("scale", C.c_double), ("events_per_base", C.c_double),
    ]


class FeatureDesc(C.Structure):
    _fields_ = [
        ("raw_pA", C.c_void_p), ("raw_dac", C.c_void_p), ("dac_offset", C.c_float), ("dac_scale", C.c_float),
        ("n_samples", C.c_uint64), ("event_start", C.c_void_p),
        ("is_reverse", C.c_int), ("ref_start", C.c_uint32), ("ref_end", C.c_uint32),
        ("called", C.c_void_p), ("n_called", C.c_uint32),
    ]


class FeatureTensors(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("signal", "core", "residual", "coords", "ref_index", "query_index", "quality")]


class ReadExtra(C.Structure):
    _fields_ = [("ref_to_query", C.c_void_p), ("is_reverse", C.c_int), ("ref_start", C.c_uint32), ("ref_end", C.c_uint32),
                ("called", C.c_void_p), ("n_called", C.c_uint32)]


READ_EXTRA_DTYPE = _np.dtype([("ref_to_query", _np.uint64), ("is_reverse", _np.int32), ("ref_start", _np.uint32),
                              ("ref_end", _np.uint32), ("called", _np.uint64), ("n_called", _np.uint32)], align=True)
assert READ_EXTRA_DTYPE.itemsize == C.sizeof(ReadExtra), (READ_EXTRA_DTYPE.itemsize, C.sizeof(ReadExtra))


class FeatureResult(C.Structure):
    _fields_ = [("status", C.c_int), ("n_pos", C.c_uint32),
                ("signal", C.POINTER(C.c_float)), ("core", C.POINTER(C.c_float)), ("residual", C.POINTER(C.c_float)),
                ("coords", C.POINTER(C.c_uint32)), ("ref_index", C.POINTER(C.c_uint32)),
                ("query_index", C.POINTER(C.c_uint32)), ("quality", C.POINTER(C.c_int32)),
                ("n_recs", C.c_uint32), ("recs", C.c_void_p)]


class AnalogueResult(C.Structure):
    _fields_ = [("status", C.c_int), ("n_sites", C.c_uint32), ("pos_on_ref", C.POINTER(C.c_uint32)),
                ("n_events", C.POINTER(C.c_uint32)), ("log_analogue", C.POINTER(C.c_double)),
                ("log_thymidine", C.POINTER(C.c_double))]


RAWDEPTH = 20
EVENTALIGN_REC_DTYPE = _np.dtype([("event", _np.uint32), ("ref_pos", _np.uint32), ("indel_score", _np.int32),
                                  ("label", _np.uint32)])


class EventT(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_float), ("mean", C.c_float), ("stdv", C.c_float),
                ("pos", C.c_int), ("state", C.c_int)]


# every symbol include/dnascent_b200.h declares
EXPORTS = [
    "dnb_default_config", "dnb_create", "dnb_destroy", "dnb_strerror", "dnb_last_error", "dnb_load_model",
    "dnb_submit", "dnb_wait", "dnb_result", "dnb_release",
    "dnb_batch_upload", "dnb_batch_run", "dnb_batch_fetch", "dnb_batch_drop_workspace", "dnb_batch_timings",
    "dnb_batch_io_bytes",
    "dnb_detect_events",
    "dnb_eexp", "dnb_eln", "dnb_lnSum", "dnb_lnProd", "dnb_lnGreaterThan", "dnb_uniformPDF", "dnb_normalPDF",
    "dnb_cauchyPDF", "dnb_sequence_probability_batch", "dnb_theil_sen_batch",
    "dnb_eventalign_batch", "dnb_eventalign_last_kernel_ms",
    "dnb_eventalign_features_batch", "dnb_features_last_kernel_ms",
    "dnb_batch_eventalign_features", "dnb_batch_feature_result", "dnb_batch_stage2_timings",
    "dnb_dorado_slice", "dnb_submit_chain",
    "dnb_expand_events", "dnb_expand_alignment", "dnb_host_register", "dnb_host_unregister", "dnb_host_alloc",
    "dnb_host_free", "dnb_trim", "dnb_batch_seg_timings", "dnb_batch_device", "dnb_host_stats", "dnb_host_phase_name",
    "dnb_batch_analogue_llr", "dnb_batch_analogue_result", "dnb_submit_llr", "dnb_batch_analogue_timings",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C dnascent_b200/csrc).  dnascent_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, d = C.c_void_p, C.c_size_t, C.c_double
    L.dnb_default_config.argtypes = [C.POINTER(Config)]
    L.dnb_default_config.restype = None
    L.dnb_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.dnb_destroy.argtypes = [vp]
    L.dnb_destroy.restype = None
    L.dnb_strerror.restype = C.c_char_p
    L.dnb_strerror.argtypes = [C.c_int]
    L.dnb_last_error.restype = C.c_char_p
    L.dnb_load_model.argtypes = [vp, C.c_int, vp, vp, sz]
    L.dnb_submit.argtypes = [vp, vp, sz, C.POINTER(vp)]
    L.dnb_batch_upload.argtypes = [vp, vp, sz, C.POINTER(vp)]
    L.dnb_wait.argtypes = [vp]
    L.dnb_batch_run.argtypes = [vp]
    L.dnb_batch_fetch.argtypes = [vp]
    L.dnb_batch_drop_workspace.argtypes = [vp]
    L.dnb_result.argtypes = [vp, sz, C.POINTER(ReadResult)]
    L.dnb_release.argtypes = [vp]
    L.dnb_release.restype = None
    L.dnb_batch_timings.argtypes = [vp, C.POINTER(d * 8), C.POINTER(C.c_uint64 * 8)]
    L.dnb_batch_io_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.dnb_detect_events.argtypes = [vp, vp, sz, C.POINTER(EventT), sz, C.POINTER(sz)]
    L.dnb_eexp.restype = d
    L.dnb_eexp.argtypes = [d]
    L.dnb_eln.argtypes = [d, C.POINTER(d)]
    for f in ("dnb_lnSum", "dnb_lnProd"):
        getattr(L, f).restype = d
        getattr(L, f).argtypes = [d, d]
    L.dnb_lnGreaterThan.argtypes = [d, d]
    for f in ("dnb_uniformPDF", "dnb_normalPDF", "dnb_cauchyPDF"):
        getattr(L, f).restype = d
        getattr(L, f).argtypes = [d, d, d]
    L.dnb_sequence_probability_batch.argtypes = [vp, vp, vp, C.c_char_p, vp, vp, vp, sz, C.c_uint32, vp, vp]
    L.dnb_theil_sen_batch.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.dnb_eventalign_batch.argtypes = [vp, vp, sz, C.c_uint32, vp, vp, vp, vp]
    L.dnb_eventalign_last_kernel_ms.restype = d
    L.dnb_eventalign_features_batch.argtypes = [vp, vp, vp, sz, C.c_uint32, vp, vp, vp, vp, vp, vp, vp]
    L.dnb_features_last_kernel_ms.restype = d
    L.dnb_batch_eventalign_features.argtypes = [vp, vp, C.c_uint32, C.c_int]
    L.dnb_submit_chain.argtypes = [vp, vp, vp, sz, C.c_uint32, C.c_int, C.POINTER(vp)]
    L.dnb_batch_feature_result.argtypes = [vp, sz, C.POINTER(FeatureResult)]
    L.dnb_dorado_slice.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64)]
    L.dnb_batch_stage2_timings.argtypes = [vp, C.POINTER(d * 2), C.POINTER(C.c_uint64 * 2)]
    L.dnb_expand_events.argtypes = [C.POINTER(ReadResult), vp]
    L.dnb_expand_alignment.argtypes = [C.POINTER(ReadResult), vp]
    L.dnb_host_register.argtypes = [vp, sz]
    L.dnb_host_unregister.argtypes = [vp]
    L.dnb_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.dnb_host_free.argtypes = [vp]
    L.dnb_host_free.restype = None
    L.dnb_trim.argtypes = [vp]
    L.dnb_batch_seg_timings.argtypes = [vp, C.POINTER(d * 3)]
    L.dnb_batch_device.argtypes = [vp]
    L.dnb_host_stats.argtypes = [C.c_int, C.POINTER(d * 17), C.POINTER(C.c_uint64 * 4)]
    L.dnb_host_phase_name.argtypes = [C.c_int]
    L.dnb_host_phase_name.restype = C.c_char_p
    L.dnb_batch_analogue_llr.argtypes = [vp, vp, C.c_uint32]
    L.dnb_batch_analogue_result.argtypes = [vp, sz, C.POINTER(AnalogueResult)]
    L.dnb_submit_llr.argtypes = [vp, vp, vp, sz, C.c_uint32, C.POINTER(vp)]
    L.dnb_batch_analogue_timings.argtypes = [vp, C.POINTER(d * 2), C.POINTER(C.c_uint64 * 4)]
    _lib = L
    return L


def check(code: int, where: str):
    if code != 0:
        L = lib()
        detail = L.dnb_strerror(code).decode()
        last = L.dnb_last_error().decode()
        raise DnbError(code, where, detail + (": " + last if last else ""))
