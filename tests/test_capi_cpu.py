"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, refuses to compute
without a GPU (no CPU fallback), and its scalar probability.h drop-ins match the reference's known answers."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from dnascent_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dnascent_b200.h")).read()
    declared = sorted(set(re.findall(r"DNB_API\s+[\w\s\*]+?\b(dnb_\w+)\s*\(", hdr)))
    assert len(declared) >= 24
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/dnascent_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_struct_layouts_match_header():
    from dnascent_b200 import _lib
    assert C.sizeof(_lib.ReadDesc) == 88 and _lib.READ_DESC_DTYPE.itemsize == 88
    assert C.sizeof(_lib.EventT) == 32
    cfg = _lib.Config()
    _lib.lib().dnb_default_config(C.byref(cfg))
    # event_detection.h:19-25 and config.h:41
    assert (cfg.window_length1, cfg.window_length2) == (3, 6)
    assert (cfg.threshold1, cfg.threshold2, cfg.peak_height) == (np.float32(1.4), np.float32(9.0), np.float32(0.2))
    assert (cfg.min_average_log_emission, cfg.max_gap_threshold, cfg.bandwidth) == (-2.0, 5, 100)


def test_ctypes_mirrors_match_the_compiled_header(tmp_path):
    """Every struct of include/dnascent_b200.h as the C compiler lays it out (sizeof + offsetof of each field, from a
    program compiled here with gcc) against the ctypes / numpy mirrors the tests and bench.py bind with."""
    import shutil
    import subprocess
    from dnascent_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    mirrors = {"dnb_config": _lib.Config, "dnb_read_desc": _lib.ReadDesc, "dnb_read_result": _lib.ReadResult,
               "dnb_event_t": _lib.EventT, "dnb_eventalign_desc": _lib.EventalignDesc, "dnb_feature_desc": _lib.FeatureDesc,
               "dnb_feature_tensors": _lib.FeatureTensors, "dnb_read_extra": _lib.ReadExtra,
               "dnb_feature_result": _lib.FeatureResult, "dnb_q2r_run": _lib.Q2RRun}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dnascent_b200.h"', 'int main(void) {']
    for cname, st in mirrors.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('printf("dnb_eventalign_rec %zu\\n", sizeof(dnb_eventalign_rec));')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, st in mirrors.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, (cname, fname)
    assert int(got["dnb_eventalign_rec"]) == _lib.EVENTALIGN_REC_DTYPE.itemsize == 16
    assert _lib.READ_DESC_DTYPE.itemsize == C.sizeof(_lib.ReadDesc) and _lib.READ_EXTRA_DTYPE.itemsize == C.sizeof(_lib.ReadExtra)
    for dt, st in ((_lib.READ_DESC_DTYPE, _lib.ReadDesc), (_lib.READ_EXTRA_DTYPE, _lib.ReadExtra)):
        for fname, _ in st._fields_:
            assert dt.fields[fname][1] == getattr(st, fname).offset, fname


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from dnascent_b200 import api
    with pytest.raises(api.DnbError) as e:
        api.Context(device=0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_probability_drop_ins_known_answers():
    from dnascent_b200 import api
    k = np.load(os.path.join(GOLDEN, "probability_v1.npz"))
    for x, e in zip(k["xs"], k["eexp"]):
        assert api.eexp(float(x)) == e
    for x, e in zip(k["xs"], k["eln"]):
        if np.isinf(e):
            with pytest.raises(api.NegativeLog):
                api.eln(float(x))
        elif np.isnan(e):
            assert math.isnan(api.eln(float(x)))
        else:
            assert api.eln(float(x)) == e
    for (a, b), s, pr, gt in zip(k["pairs"], k["lnSum"], k["lnProd"], k["lnGreaterThan"]):
        for got, want in ((api.lnSum(float(a), float(b)), s), (api.lnProd(float(a), float(b)), pr)):
            assert (math.isnan(got) and math.isnan(want)) or got == want
        assert api.lnGreaterThan(float(a), float(b)) == bool(gt)
    for t, u, n, c in zip(k["triples"], k["uniformPDF"], k["normalPDF"], k["cauchyPDF"]):
        t = [float(v) for v in t]
        assert api.uniformPDF(*t) == u and api.normalPDF(*t) == n and api.cauchyPDF(*t) == c


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under dnascent_b200/ may import or link it."""
    pkg = os.path.join(ROOT, "dnascent_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "dnb_oracle" not in text, f


def test_shim_exports_the_reference_symbols():
    """The C++ drop-in layer defines, with the reference's exact (mangled) signatures, every symbol of the hot path
    that detect.cpp / alignment.cpp / trainCNN.cpp link against (event_handling.h:13, event_detection.h:35,
    probability.h:26-33, detect.h:119,121).  Built only where the reference headers are mounted."""
    import subprocess
    from oracle import refbind
    if not refbind.shim_available():
        pytest.skip("oracle/_ref/libdnascent_shim.so not built (needs /root/reference at build time)")
    out = subprocess.run(["nm", "-D", "--defined-only", refbind.SHIM_LIB_PATH], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for sym in ("_Z15normaliseEventsRN8DNAscent4readEb", "detect_events", "_Z4eexpd", "_Z3elnd", "_Z5lnSumdd", "_Z6lnProddd",
                "_Z13lnGreaterThandd", "_Z10uniformPDFddd", "_Z9normalPDFddd", "_Z9cauchyPDFddd",
                "_Z12llAcrossReadRN8DNAscent4readEj", "_Z10eventalignRN8DNAscent4readEj"):
        assert sym in defined, sym
    assert any(s.startswith("_Z19sequenceProbabilityRSt6vectorIdSaIdEE") for s in defined)
    refbind.Ref(shim=True)          # loads (resolves libdnascent_b200.so through its rpath) without a GPU


def test_dorado_slice_matches_vector_erase():
    """Row f3: dnb_dorado_slice against a literal restatement of pod5_getSignal's two vector::erase calls
    (src/pod5.cpp:76-93) on a Python list, for unsplit and split reads; undefined cases are refused."""
    from dnascent_b200 import _lib, api
    rng = np.random.default_rng(5)

    def erase_twice(raw, sig_start, sig_end):                 # pod5.cpp:84-85 / 92-93
        raw = list(raw)
        del raw[:sig_start]
        del raw[sig_end - sig_start:]
        return raw

    for _ in range(200):
        n = int(rng.integers(1, 400))
        raw = list(range(n))
        split = bool(rng.integers(0, 2))
        sp = int(rng.integers(0, n)) if split else int(rng.integers(0, 1000))   # sp is ignored unless the read was split
        lo = sp if split else 0
        ns = int(rng.integers(1, n - lo + 1))
        ts = int(rng.integers(0, ns + 1))
        sl = api.dorado_slice(n, ns, ts, sp, split)
        want = erase_twice(raw, lo + ts, lo + ns)
        assert raw[sl] == want
    assert api.dorado_slice(100) == slice(0, 100)              # no ns tag: signalLength stays 0, nothing is trimmed
    assert api.dorado_slice(100, 100, 100) == slice(100, 100)  # everything trimmed: empty read (fails later, as in the reference)
    for bad in ((100, 101, 0, 0, False), (100, 50, 51, 0, False), (100, 60, 0, 50, True), (0, 0, 0, 0, False),
                (100, 50, -1, 0, False)):
        with pytest.raises(_lib.DnbError):
            api.dorado_slice(*bad)


def test_compact_result_expanders_are_the_inverse_of_the_wire_format():
    """dnb_expand_events / dnb_expand_alignment are host code (no device needed): hand-built compact results, including
    escaped event lengths and every step kind, must expand to the dense arrays they were made from."""
    from dnascent_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for trial in range(20):
        ne = int(rng.integers(1, 300))
        lens = rng.integers(1, 40, size=ne).astype(np.uint32)
        lens[rng.random(ne) < 0.05] = rng.integers(255, 100000, size=int((rng.random(ne) < 0.05).sum()) or 1)[0]
        first = int(rng.integers(0, 5))
        starts = np.concatenate([[first], first + np.cumsum(lens.astype(np.uint64))]).astype(np.uint32)
        len8 = np.minimum(lens, 255).astype(np.uint8)
        esc = np.ascontiguousarray(lens[lens >= 255], dtype=np.uint32)
        na = int(rng.integers(1, 500))
        codes = rng.integers(0, 3, size=na - 1).astype(np.uint8)
        e0, k0 = int(rng.integers(0, 10)), int(rng.integers(0, 10))
        ev = e0 + np.concatenate([[0], np.cumsum(codes != 2)])
        km = k0 + np.concatenate([[0], np.cumsum(codes != 1)])
        packed = np.zeros((na - 1 + 3) // 4 + 1, dtype=np.uint8)
        for t, c in enumerate(codes):
            packed[t >> 2] |= int(c) << (2 * (t & 3))
        r = _lib.ReadResult()
        r.n_events, r.n_align = ne, na
        r.event_first = first
        r.event_len8 = len8.ctypes.data_as(C.POINTER(C.c_uint8))
        r.event_len_escape = esc.ctypes.data_as(C.POINTER(C.c_uint32)) if esc.size else None
        r.n_event_len_escape = esc.size
        r.align_first[0], r.align_first[1] = e0, k0
        r.align_steps = packed.ctypes.data_as(C.POINTER(C.c_uint8))
        out_s = np.zeros(ne + 1, dtype=np.uint32)
        out_p = np.zeros((na, 2), dtype=np.uint32)
        assert L.dnb_expand_events(C.byref(r), out_s.ctypes.data) == 0
        assert L.dnb_expand_alignment(C.byref(r), out_p.ctypes.data) == 0
        np.testing.assert_array_equal(out_s, starts)
        np.testing.assert_array_equal(out_p[:, 0], ev)
        np.testing.assert_array_equal(out_p[:, 1], km)
        if esc.size:       # a truncated escape list is reported, not read past
            r.n_event_len_escape = esc.size - 1
            assert L.dnb_expand_events(C.byref(r), out_s.ctypes.data) == 5
