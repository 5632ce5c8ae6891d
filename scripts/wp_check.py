"""One-shot check of the experimental window-parallel eventalign (DNB_EA_WINDOW_PARALLEL=1) against the reference's golden
records and against the read-serial kernel's tensors.  Run on a GPU box: DNB_EA_WINDOW_PARALLEL=1 python scripts/wp_check.py"""
import os, sys, time
t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import conftest
from test_eventalign_cpu import all_golden_reads, golden_eventalign_inputs
from dnascent_b200 import api
e = np.load(os.path.join(conftest.GOLDEN, "eventalign_v1.npz"))
reads = all_golden_reads(conftest.golden_reads.__wrapped__(), conftest.golden_v2.__wrapped__())
mean = np.load(os.path.join(conftest.GOLDEN, "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)
ctx = api.Context(0)
ctx.load_model(api.MODEL_PORE, mean)
print("setup %.1f s, window-parallel=%s" % (time.time() - t0, os.environ.get("DNB_EA_WINDOW_PARALLEL")), flush=True)
out = ctx.eventalign([golden_eventalign_inputs(g, e, tag) for tag, g in reads], window=50)
bad = 0
for (tag, g), rec in zip(reads, out):
    same = rec["status"] == 0 and all(np.array_equal(rec[k], e[f"e_{tag}_{k}"]) for k in ("event", "ref_pos", "label", "indel"))
    print(tag, "status", rec["status"], "records", rec["event"].size, "golden", e[f"e_{tag}_event"].size, "IDENTICAL" if same else "DIFFERENT", flush=True)
    bad += not same
print("kernel ms", ctx.eventalign_last_kernel_ms(), "total %.1f s" % (time.time() - t0), "FAILED" if bad else "ALL IDENTICAL")
