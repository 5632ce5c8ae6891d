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


@pytest.fixture(scope="session")
def ctx(pore_mean):
    from dnascent_b200 import api
    c = api.Context(device=0, keep_debug=True)
    c.load_model(api.MODEL_PORE, pore_mean)
    yield c
    c.close()
