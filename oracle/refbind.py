"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libdnascent_ref.so (the UNMODIFIED reference
sources, built by oracle/Makefile `ref`).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libdnascent_ref.so")
# same harness + reference objects, but the hot-path symbols come from the product's C++ shim over the CUDA library
SHIM_LIB_PATH = os.path.join(HERE, "_ref", "libdnascent_shim.so")

PORE, UNLABELLED, ANALOGUE = 0, 1, 2
NKMER = 4 ** 9


def available() -> bool:
    return os.path.exists(LIB_PATH)


def shim_available() -> bool:
    return os.path.exists(SHIM_LIB_PATH)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Ref:
    """One process-wide instance (the reference keeps its tables in a global)."""

    def __init__(self, shim: bool = False):
        path = SHIM_LIB_PATH if shim else LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle {'shim' if shim else 'ref'}` where /root/reference exists")
        self.is_shim = shim
        L = self.L = C.CDLL(path)
        sz, vp, d = C.c_size_t, C.c_void_p, C.c_double
        if shim:
            L.dnbshim_normalise_batch.argtypes = [vp, sz]
            L.dnbshim_ll_across_read_batch.argtypes = [vp, sz, C.c_uint]
            L.dnbshim_calls.restype = sz
            L.dnbshim_calls.argtypes = [vp, vp, vp, sz]
            L.dnbshim_shutdown.restype = None
            L.dnbshim_eventalign_features_batch.argtypes = [vp, sz, C.c_uint]
            L.dnbshim_normalise_eventalign_batch.argtypes = [vp, sz, C.c_uint]
            L.dnbshim_dnn_inputs.restype = sz
            L.dnbshim_dnn_inputs.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp, sz]
        L.dnbref_get_model.restype = sz
        L.dnbref_get_model.argtypes = [C.c_int, vp, vp, sz]
        L.dnbref_set_model.argtypes = [C.c_int, vp, vp, sz]
        L.dnbref_set_reference.argtypes = [C.c_char_p, C.c_char_p, sz]
        L.dnbref_read_new.restype = vp
        L.dnbref_read_new.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, vp, C.c_uint32, C.c_int, C.c_int32, vp, sz]
        L.dnbref_read_free.argtypes = [vp]
        for f in ("dnbref_read_basecall", "dnbref_read_refseq", "dnbref_read_query_to_ref", "dnbref_read_ref_to_query"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp, vp, sz]
        for f in ("dnbref_read_is_reverse", "dnbref_read_ref_start", "dnbref_read_ref_end"):
            getattr(L, f).argtypes = [vp]
        L.dnbref_normalise.argtypes = [vp, C.c_int]
        if not shim:
            L.dnbref_normalise_staged.argtypes = [vp, C.c_int]
        L.dnbref_n_events.restype = sz
        L.dnbref_n_events.argtypes = [vp]
        L.dnbref_events.restype = sz
        L.dnbref_events.argtypes = [vp, vp, vp, sz]
        L.dnbref_events_raw_concat.restype = sz
        L.dnbref_events_raw_concat.argtypes = [vp, vp, sz]
        L.dnbref_alignment.restype = sz
        L.dnbref_alignment.argtypes = [vp, vp, vp, sz]
        L.dnbref_scalars.argtypes = [vp, vp]
        L.dnbref_cleaned.restype = sz
        L.dnbref_cleaned.argtypes = [vp, vp, vp, sz]
        L.dnbref_detect_events.restype = sz
        L.dnbref_detect_events.argtypes = [vp, sz, vp, vp, vp, vp, sz]
        if not shim:
            L.dnbref_theil_sen.argtypes = [vp, vp, sz, d, d, C.c_int, vp, vp]
        L.dnbref_sequence_probability.restype = d
        L.dnbref_sequence_probability.argtypes = [vp, sz, C.c_char_p, sz, C.c_int, d, d, d, sz, sz]
        L.dnbref_ll_across_read.restype = sz
        L.dnbref_ll_across_read.argtypes = [vp, C.c_uint, vp, vp, sz]
        L.dnbref_kmer2index.restype = C.c_uint
        L.dnbref_kmer2index.argtypes = [C.c_char_p, C.c_uint]
        for f in ("dnbref_eexp",):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d]
        L.dnbref_eln.argtypes = [d, vp]
        for f in ("dnbref_lnSum", "dnbref_lnProd"):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d, d]
        L.dnbref_lnGreaterThan.argtypes = [d, d]
        for f in ("dnbref_uniformPDF", "dnbref_normalPDF", "dnbref_cauchyPDF"):
            getattr(L, f).restype = d
            getattr(L, f).argtypes = [d, d, d]
        L.dnbref_eventalign.restype = sz
        L.dnbref_eventalign.argtypes = [vp, C.c_uint, vp, sz]
        L.dnbref_aligned_positions.restype = sz
        L.dnbref_aligned_positions.argtypes = [vp, vp, vp, vp, vp, sz]
        L.dnbref_aligned_indices.restype = sz
        L.dnbref_aligned_indices.argtypes = [vp, vp, vp, vp, sz]
        L.dnbref_rawdepth.restype = sz
        if not shim:
            L.dnbref_builtin_viterbi.restype = sz
            L.dnbref_builtin_viterbi.argtypes = [vp, sz, C.c_char_p, d, d, d, vp, vp, vp, sz]
        L.dnbref_bench_normalise.restype = d
        L.dnbref_bench_normalise.argtypes = [vp, sz, C.c_int, C.c_int, vp]

    # -- configuration ------------------------------------------------------------------
    def configure_from_files(self):
        if self.L.dnbref_configure_from_files() != 0:
            raise RuntimeError("reference model files not readable")

    def load_model_file(self, which: int, filename: str, fit_stdv: bool = True):
        """One table through the reference's own parser, e.g. ANALOGUE <- r10.4.1_EdU_gaussian.model (build container only)."""
        self.L.dnbref_load_model_file.argtypes = [C.c_int, C.c_char_p, C.c_int]
        if self.L.dnbref_load_model_file(which, filename.encode(), int(fit_stdv)) != 0:
            raise RuntimeError(f"reference model file {filename} not readable")

    def set_model(self, which: int, mean: np.ndarray, stdv: np.ndarray):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        stdv = np.ascontiguousarray(stdv, dtype=np.float64)
        self.L.dnbref_set_model(which, _p(mean), _p(stdv), mean.size)

    def get_model(self, which: int):
        m = np.zeros(NKMER)
        s = np.zeros(NKMER)
        n = self.L.dnbref_get_model(which, _p(m), _p(s), NKMER)
        return m[:n], s[:n]

    def set_reference(self, seq: bytes, name: bytes = b"chr1"):
        self.L.dnbref_set_reference(name, seq, len(seq))

    def kmer2index(self, kmer: bytes) -> int:
        return self.L.dnbref_kmer2index(kmer, len(kmer))

    # -- reads --------------------------------------------------------------------------
    def read_new(self, sr) -> "RefRead":
        raw = np.ascontiguousarray(sr.raw, dtype=np.float32)
        cig = np.ascontiguousarray(sr.cigar, dtype=np.uint32)
        h = self.L.dnbref_read_new(sr.name.encode(), sr.seq_bam, len(sr.seq_bam), _p(cig), cig.size, sr.flag, sr.pos,
                                   _p(raw), raw.size)
        if not h:
            raise RuntimeError("reference read constructor failed")
        return RefRead(self, h)

    def detect_events(self, raw: np.ndarray):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        cap = raw.size + 2
        start = np.zeros(cap, dtype=np.uint64)
        length = np.zeros(cap, dtype=np.float32)
        mean = np.zeros(cap, dtype=np.float32)
        stdv = np.zeros(cap, dtype=np.float32)
        n = self.L.dnbref_detect_events(_p(raw), raw.size, _p(start), _p(length), _p(mean), _p(stdv), cap)
        return start[:n], length[:n], mean[:n], stdv[:n]

    def theil_sen(self, sig, ranks, shift, scale, use_fit=False):
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
        o = np.zeros(2)
        self.L.dnbref_theil_sen(_p(sig), _p(ranks), sig.size, shift, scale, int(use_fit), _p(o[0:1]), _p(o[1:2]))
        return float(o[0]), float(o[1])

    def sequence_probability(self, obs, seq: bytes, window: int, use_brdu: bool, shift, scale, events_per_base,
                             start: int, end: int) -> float:
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        return self.L.dnbref_sequence_probability(_p(obs), obs.size, seq, window, int(use_brdu), shift, scale,
                                                  events_per_base, start, end)

    def builtin_viterbi(self, obs, seq: bytes, shift, scale, events_per_base):
        """builtinViterbi (alignment.cpp:193-516) on one window: (score, state index[], state type[] 0=D 1=M 2=I)."""
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        cap = 4 * (obs.size + len(seq)) + 16
        idx = np.zeros(cap, dtype=np.int32)
        typ = np.zeros(cap, dtype=np.uint8)
        score = C.c_double(0.0)
        n = self.L.dnbref_builtin_viterbi(_p(obs), obs.size, seq, shift, scale, events_per_base, C.byref(score),
                                          _p(idx), _p(typ), cap)
        return score.value, idx[:n].copy(), typ[:n].copy()

    def eln(self, x: float):
        o = np.zeros(1)
        if self.L.dnbref_eln(x, _p(o)):
            raise ValueError("NegativeLog")
        return float(o[0])

    # -- shim build only: the batched entry points ---------------------------------------
    def normalise_batch(self, reads):
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        self.L.dnbshim_normalise_batch(arr, len(reads))

    def ll_across_read_batch(self, reads, window: int = 12):
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        self.L.dnbshim_ll_across_read_batch(arr, len(reads), window)
        out = []
        for r in reads:
            cap = len(r.refseq) + 1
            pos = np.zeros(cap, dtype=np.int32)
            llr = np.zeros(cap)
            n = self.L.dnbshim_calls(r.h, _p(pos), _p(llr), cap)
            out.append((pos[:n].copy(), llr[:n].copy()))
        return out

    def normalise_ll_batch(self, reads, window: int = 12):
        """dnb_shim::normalise_llAcrossRead_batch on reads that have NOT been normalised yet; returns the calls per read."""
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        self.L.dnbshim_normalise_ll_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
        self.L.dnbshim_normalise_ll_batch(arr, len(reads), window)
        out = []
        for r in reads:
            cap = len(r.refseq) + 1
            pos = np.zeros(cap, dtype=np.int32)
            llr = np.zeros(cap)
            n = self.L.dnbshim_calls(r.h, _p(pos), _p(llr), cap)
            out.append((pos[:n].copy(), llr[:n].copy()))
        return out

    def eventalign_features_batch(self, reads, window: int = 50, resident: bool = False):
        """dnb_shim::eventalign_features_batch (or, resident=True, dnb_shim::normalise_eventalign_batch on reads that have
        NOT been normalised yet): per read the DnnInputs vectors (what runCNN would get from r.makeSignalTensor() & co.),
        keys as aligned_positions()."""
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        fn = self.L.dnbshim_normalise_eventalign_batch if resident else self.L.dnbshim_eventalign_features_batch
        if fn(arr, len(reads), window):
            raise ValueError("NegativeLog")
        depth = self.L.dnbref_rawdepth()
        out = []
        for i, r in enumerate(reads):
            cap = len(r.refseq) + 1
            sig = np.zeros(cap * depth, dtype=np.float32)
            f32 = [np.zeros(cap, dtype=np.float32) for _ in range(2)]
            u32 = [np.zeros(cap, dtype=np.uint32) for _ in range(3)]
            q = np.zeros(cap, dtype=np.int32)
            P = self.L.dnbshim_dnn_inputs(i, _p(sig), _p(f32[0]), _p(f32[1]), _p(u32[0]), _p(u32[1]), _p(u32[2]), _p(q), cap)
            assert P <= cap
            out.append(dict(signal=sig[:P * depth].reshape(P, depth).copy(), core=f32[0][:P].copy(),
                            residual=f32[1][:P].copy(), coords=u32[0][:P].copy(), ref_index=u32[1][:P].copy(),
                            query_index=u32[2][:P].copy(), quality=q[:P].copy()))
        return out

    def shutdown(self):
        if self.is_shim:
            self.L.dnbshim_shutdown()

    def set_devices(self, devices):
        """dnb_shim::set_devices: the one process drives all of them (call after shutdown(), before the next batch)."""
        arr = (C.c_int * len(devices))(*devices)
        self.L.dnbshim_set_devices(arr, len(devices))

    def batches_on_device(self, device: int) -> int:
        self.L.dnbshim_batches_on_device.restype = C.c_ulong
        return int(self.L.dnbshim_batches_on_device(device))

    def bench_chain(self, reads, threads: int, window: int = 50):
        """normaliseEvents + eventalign + tensor builders per read on `threads` host threads (detect.cpp:876-888)."""
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        failed = C.c_int(0)
        self.L.dnbref_bench_chain.restype = C.c_double
        self.L.dnbref_bench_chain.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.POINTER(C.c_int)]
        t = self.L.dnbref_bench_chain(arr, len(reads), threads, window, C.byref(failed))
        return t, failed.value

    def bench_hmm(self, reads, threads: int, window: int = 12):
        """normaliseEvents + llAcrossRead per read on `threads` host threads (detect --HMM, detect.cpp:876-885);
        returns (wall seconds, failed reads, LLR calls made)."""
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        failed, calls = C.c_int(0), C.c_longlong(0)
        self.L.dnbref_bench_hmm.restype = C.c_double
        self.L.dnbref_bench_hmm.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
        t = self.L.dnbref_bench_hmm(arr, len(reads), threads, window, C.byref(failed), C.byref(calls))
        return t, failed.value, calls.value

    def bench_normalise(self, reads, threads: int, use_fit=False):
        arr = (C.c_void_p * len(reads))(*[r.h for r in reads])
        failed = C.c_int(0)
        t = self.L.dnbref_bench_normalise(arr, len(reads), threads, int(use_fit), C.byref(failed))
        return t, failed.value


class RefRead:
    def __init__(self, ref: Ref, h):
        self.ref, self.h, self.L = ref, h, ref.L

    def free(self):
        if self.h:
            self.L.dnbref_read_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _str(self, fn):
        n = fn(self.h, None, 0)
        b = C.create_string_buffer(n)
        fn(self.h, b, n)
        return b.raw[:n]

    @property
    def basecall(self) -> bytes:
        return self._str(self.L.dnbref_read_basecall)

    @property
    def refseq(self) -> bytes:
        return self._str(self.L.dnbref_read_refseq)

    @property
    def query_to_ref(self) -> np.ndarray:
        n = len(self.basecall)
        a = np.zeros(n, dtype=np.int32)
        self.L.dnbref_read_query_to_ref(self.h, _p(a), n)
        return a

    @property
    def ref_to_query(self) -> np.ndarray:
        n = len(self.refseq)
        a = np.zeros(n, dtype=np.int32)
        self.L.dnbref_read_ref_to_query(self.h, _p(a), n)
        return a

    def normalise(self, staged=True, use_fit=False) -> dict:
        """Run the reference normaliseEvents (staged: also re-derive per-stage values) and collect outputs."""
        if staged:
            rc = self.L.dnbref_normalise_staged(self.h, int(use_fit))
            if rc != 0:
                raise RuntimeError("staged re-derivation disagrees with normaliseEvents")
        else:
            self.L.dnbref_normalise(self.h, int(use_fit))
        return self.outputs(staged)

    def outputs(self, staged=True) -> dict:
        ne = self.L.dnbref_n_events(self.h)
        means = np.zeros(ne)
        rlen = np.zeros(ne, dtype=np.uint32)
        self.L.dnbref_events(self.h, _p(means), _p(rlen), ne)
        cap = ne * 2 + len(self.basecall) + 16
        ev = np.zeros(cap, dtype=np.uint32)
        km = np.zeros(cap, dtype=np.uint32)
        na = self.L.dnbref_alignment(self.h, _p(ev), _p(km), cap)
        sc = np.zeros(9)
        self.L.dnbref_scalars(self.h, _p(sc))
        out = dict(event_mean=means, event_raw_len=rlen, align_event=ev[:na].copy(), align_kmer=km[:na].copy(),
                   shift=sc[0], scale=sc[1], events_per_base=sc[2], avg_log_emission=sc[5], spanned=bool(sc[6]),
                   max_gap=int(sc[7]), qc_set=bool(sc[8]))
        if staged:
            csig = np.zeros(cap)
            crk = np.zeros(cap, dtype=np.uint32)
            nc = self.L.dnbref_cleaned(self.h, _p(csig), _p(crk), cap)
            out.update(rough_shift=sc[3], rough_scale=sc[4], cleaned_signal=csig[:nc].copy(),
                       cleaned_rank=crk[:nc].copy())
        return out

    @property
    def is_reverse(self) -> bool:
        return bool(self.L.dnbref_read_is_reverse(self.h))

    @property
    def ref_start(self) -> int:
        return self.L.dnbref_read_ref_start(self.h)

    @property
    def ref_end(self) -> int:
        return self.L.dnbref_read_ref_end(self.h)

    def events_raw_concat(self) -> np.ndarray:
        """The raw pA slices of r.events, concatenated (lengths: outputs()['event_raw_len'])."""
        n = self.L.dnbref_events_raw_concat(self.h, None, 0)
        a = np.zeros(n)
        self.L.dnbref_events_raw_concat(self.h, _p(a), n)
        return a

    def eventalign(self, window: int = 50) -> bytes:
        """eventalign (alignment.cpp:547-744) on a normalised read: humanReadable_eventalignOut."""
        n = self.L.dnbref_eventalign(self.h, window, None, 0)
        b = C.create_string_buffer(n + 1)
        self.L.dnbref_eventalign(self.h, window, b, n)
        return b.raw[:n]

    def aligned_positions(self) -> dict:
        """What eventalign left in refCoordToAP, as the DNN tensor builders (reads.h:305-372) see it."""
        P = self.L.dnbref_aligned_positions(self.h, None, None, None, None, 0)
        depth = self.L.dnbref_rawdepth()
        sig = np.zeros(P * depth, dtype=np.float32)
        core = np.zeros(P, dtype=np.float32)
        resid = np.zeros(P, dtype=np.float32)
        coords = np.zeros(P, dtype=np.uint32)
        if P:
            self.L.dnbref_aligned_positions(self.h, _p(sig), _p(core), _p(resid), _p(coords), P)
        ref_index = np.zeros(P, dtype=np.uint32)
        query_index = np.zeros(P, dtype=np.uint32)
        quality = np.zeros(P, dtype=np.int32)
        if P:
            self.L.dnbref_aligned_indices(self.h, _p(ref_index), _p(query_index), _p(quality), P)
        return dict(signal=sig.reshape(P, depth), core=core, residual=resid, coords=coords, ref_index=ref_index,
                    query_index=query_index, quality=quality)

    def ll_across_read(self, window: int = 12):
        cap = len(self.refseq) + 1
        pos = np.zeros(cap, dtype=np.int32)
        llr = np.zeros(cap)
        n = self.L.dnbref_ll_across_read(self.h, window, _p(pos), _p(llr), cap)
        return pos[:n].copy(), llr[:n].copy()
