"""Timing of eventalign on the golden reads replicated (read-serial vs DNB_EA_WINDOW_PARALLEL=1, one process each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import conftest
from test_eventalign_cpu import all_golden_reads, golden_eventalign_inputs
from dnascent_b200 import api
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 100
e = np.load(os.path.join(conftest.GOLDEN, "eventalign_v1.npz"))
reads = all_golden_reads(conftest.golden_reads.__wrapped__(), conftest.golden_v2.__wrapped__())
mean = np.load(os.path.join(conftest.GOLDEN, "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)
ctx = api.Context(0)
ctx.load_model(api.MODEL_PORE, mean)
ins = [golden_eventalign_inputs(g, e, tag) for tag, g in reads] * rep
best = 1e30
for _ in range(3):
    out = ctx.eventalign(ins, window=50)
    best = min(best, ctx.eventalign_last_kernel_ms())
ok = all(np.array_equal(out[i]["event"], e[f"e_{reads[i % len(reads)][0]}_event"]) for i in range(len(out)))
print("window-parallel" if os.environ.get("DNB_EA_WINDOW_PARALLEL") == "1" else "read-serial", len(ins), "reads", "%.2f ms" % best, "records ok" if ok else "RECORDS DIFFER")
