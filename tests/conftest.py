import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pore_mean():
    return np.load(os.path.join(GOLDEN, "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)


class GoldenRead:
    def __init__(self, d, i):
        from dnascent_b200 import synth
        p = f"r{i}_"
        self.index = i
        self.seq_bam = d[p + "seq_bam"].tobytes()
        self.flag = int(d[p + "flag"])
        self.pos = int(d[p + "pos"])
        self.cigar = d[p + "cigar"]
        self.dac = d[p + "dac"]
        self.raw = synth.dac_to_pa(self.dac)
        self.basecall = d[p + "basecall"].tobytes()
        self.refseq = d[p + "refseq"].tobytes()
        self.query_to_ref = d[p + "query_to_ref"]
        self.name = f"g{i}"
        self.et_start = d[p + "et_start"]
        self.et_length = d[p + "et_length"]
        self.et_mean = d[p + "et_mean"]
        self.et_stdv = d[p + "et_stdv"]
        self.event_mean = d[p + "event_mean"]
        self.event_raw_len = d[p + "event_raw_len"]
        self.align = d[p + "align"]
        (self.shift, self.scale, self.events_per_base, self.rough_shift, self.rough_scale, self.avg_log_emission,
         sp, mg) = d[p + "scalars"]
        self.spanned, self.max_gap = bool(sp), int(mg)
        self.cleaned_signal = d[p + "cleaned_signal"]
        self.cleaned_rank = d[p + "cleaned_rank"]


@pytest.fixture(scope="session")
def golden_reads():
    d = np.load(os.path.join(GOLDEN, "reads_v1.npz"))
    return [GoldenRead(d, i) for i in range(int(d["n_reads"]))]


class GoldenReadV2(GoldenRead):
    """reads_v2.npz: analogue-substituted reads (a*) and reads with indel / soft-clip CIGARs (i*)."""

    def __init__(self, d, tag):
        from dnascent_b200 import synth
        p = tag + "_"
        self.index = tag
        self.name = tag
        self.seq_bam = d[p + "seq_bam"].tobytes()
        self.flag = int(d[p + "flag"])
        self.pos = int(d[p + "pos"])
        self.cigar = d[p + "cigar"]
        self.dac = d[p + "dac"]
        self.raw = synth.dac_to_pa(self.dac)
        self.basecall = d[p + "basecall"].tobytes()
        self.refseq = d[p + "refseq"].tobytes()
        self.query_to_ref = d[p + "query_to_ref"]
        self.et_n = int(d[p + "et_n"])
        self.event_mean = d[p + "event_mean"]
        self.event_raw_len = d[p + "event_raw_len"]
        self.align = d[p + "align"]
        (self.shift, self.scale, self.events_per_base, self.rough_shift, self.rough_scale, self.avg_log_emission,
         sp, mg) = d[p + "scalars"]
        self.spanned, self.max_gap = bool(sp), int(mg)
        self.cleaned_signal = d[p + "cleaned_signal"]
        self.cleaned_rank = d[p + "cleaned_rank"]
        if p + "llr" in d:
            self.pos_global, self.llr = d[p + "pos_global"], d[p + "llr"]


@pytest.fixture(scope="session")
def golden_v2():
    d = np.load(os.path.join(GOLDEN, "reads_v2.npz"))
    reads = {t: GoldenReadV2(d, t) for t in ("a0", "a1", "i0", "i1")}
    tables = []
    for name in ("unl_mean", "unl_stdv", "ana_mean", "ana_stdv"):
        t = np.zeros(4 ** 9)
        t[d["ranks"]] = d[name]
        tables.append(t)
    return reads, tables, d["reference"].tobytes()


@pytest.fixture(scope="session")
def golden_reference():
    return np.load(os.path.join(GOLDEN, "reads_v1.npz"))["reference"].tobytes()


@pytest.fixture(scope="session")
def port():
    from oracle import portbind
    return portbind.Port()


@pytest.fixture(scope="session")
def ref_oracle(pore_mean):
    """oracle/_ref: the unmodified reference.  Present in the build container (and shipped, prebuilt, to the GPU box)."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = refbind.Ref()
    R.set_model(refbind.PORE, pore_mean, np.full(pore_mean.size, 0.14))
    return R


def cuda_device_count() -> int:
    """Devices the library can use, asked through the library itself (0 on a CPU-only machine)."""
    from dnascent_b200 import api, _lib
    try:
        c = api.Context(device=0)
    except _lib.DnbError as ex:
        if ex.code == 2:          # DNB_ERR_CUDA: no device -- the library has no CPU path
            return 0
        raise
    c.close()
    try:
        import torch
        return max(int(torch.cuda.device_count()), 1)
    except Exception:  # noqa: BLE001
        return 1


@pytest.fixture(scope="session")
def n_cuda():
    return cuda_device_count()


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not errored) on a machine without a CUDA device: the shim's entry points abort()
    the interpreter when dnb_create fails, so they must never be reached there."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    if cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libdnascent_b200 has no CPU fallback)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def ctx(pore_mean):
    from dnascent_b200 import api
    c = api.Context(device=0, keep_debug=True)
    c.load_model(api.MODEL_PORE, pore_mean)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_compact(pore_mean):
    """The same library with the compact wire format (dnb_config.result_format = DNB_RESULT_COMPACT)."""
    from dnascent_b200 import api
    c = api.Context(device=0, keep_debug=True, result_format=api.RESULT_COMPACT)
    c.load_model(api.MODEL_PORE, pore_mean)
    yield c
    c.close()
