"""Calibration run: C1-shaped batch (n reads x L bases), device pipeline timings per stage."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnascent_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mean = np.load("tests/golden/pore_model_r10.4.1_400bps.npz")["mean"].astype(np.float64)
ref = synth.make_reference(1_000_000, 1)
t = time.time()
base = synth.simulate_batch(ref, [L] * n, mean, seed=2)
print(f"generated {n} reads in {time.time()-t:.1f}s", flush=True)
reads = [api.Read.from_synth(r, use_dac=True) for r in base] * rep
ctx = api.Context(0)
ctx.load_model(api.MODEL_PORE, mean)
t = time.time(); b = ctx.upload(reads); print(f"upload {time.time()-t:.2f}s")
for it in range(3):
    t = time.time(); b.run(); dt = time.time() - t
    ms, cnt = b.timings()
    print(f"run {it}: wall {dt*1e3:.1f} ms", {k: round(v, 2) for k, v in ms.items()}, cnt)
    print(f"   -> {cnt['samples']/ms['total']/1e3:.1f} Msamples/s device; DP {cnt['cells']/ms['banded_dp']/1e6:.2f} Gcells/s; seg {cnt['samples']/ms['segmentation']/1e3:.1f} Msamples/s")
t = time.time(); b.fetch(); print(f"fetch {time.time()-t:.2f}s")
res = b.results()
print("status counts", np.bincount([r.status for r in res], minlength=5))
b.release()
t = time.time(); out = ctx.normaliseEvents(reads); print(f"e2e submit+wait+results {time.time()-t:.2f}s")
