"""CPU tests of the host-side logic: the synthetic generator, length-balanced sharding/binning and the
world_size-2 shard plan (gloo)."""
import os
import sys

import numpy as np
import pytest

from dnascent_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_pod5_shaped(pore_mean):
    ref = synth.make_reference(30_000, 5)
    a = synth.simulate_batch(ref, [1500, 2500], pore_mean, seed=9)
    b = synth.simulate_batch(ref, [1500, 2500], pore_mean, seed=9)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x.dac, y.dac)
        assert x.basecall == y.basecall and x.flag == y.flag
    r = a[0]
    # float32-exact pA, the expression shape of pod5.cpp:60
    np.testing.assert_array_equal(r.raw, (r.dac.astype(np.float32) + np.float32(-240.0)) * np.float32(0.1465))
    assert r.raw.dtype == np.float32 and 8 < r.raw.size / len(r.basecall) < 16
    if r.flag == 16:
        assert r.basecall == synth.revcomp(r.seq_bam)
    np.testing.assert_array_equal(r.query_to_ref, np.arange(len(r.basecall)))


def test_lognormal_n50():
    L = synth.lognormal_lengths(200_000, 30_000, np.random.default_rng(1))
    s = np.sort(L)[::-1]
    n50 = s[np.searchsorted(np.cumsum(s), s.sum() / 2)]
    assert abs(n50 - 30_000) / 30_000 < 0.03


def test_shard_reads_balanced_and_complete():
    rng = np.random.default_rng(3)
    n = synth.lognormal_lengths(20_000, 30_000, rng) * 12
    for ranks in (1, 2, 4, 8):
        sh = sharding.shard_reads(n, ranks)
        allidx = np.sort(np.concatenate(sh))
        np.testing.assert_array_equal(allidx, np.arange(n.size))           # every read exactly once
        loads = np.array([n[s].sum() for s in sh], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01                           # length-balanced
    assert sharding.shard_reads([], 2)[0].size == 0


def test_make_bins_respects_budget_and_orders_longest_first():
    rng = np.random.default_rng(4)
    n = synth.lognormal_lengths(5_000, 30_000, rng) * 12
    for policy in ("interleaved", "sorted"):
        bins = sharding.make_bins(n, 20_000_000, policy=policy)
        np.testing.assert_array_equal(np.sort(np.concatenate(bins)), np.arange(n.size))
        for b in bins:
            assert n[b].sum() <= 20_000_000 or b.size == 1
            assert np.all(np.diff(n[b]) <= 0)                                  # longest first inside a bin
        firsts = [n[b[0]] for b in bins]
        assert firsts == sorted(firsts, reverse=True)
    # "interleaved" gives every bin the same length mix: equal loads, and every bin holds some of the longest reads
    bins = sharding.make_bins(n, 20_000_000, policy="interleaved")
    loads = np.array([n[b].sum() for b in bins], dtype=np.float64)
    assert loads.max() / loads.min() < 1.15
    assert min(n[b[0]] for b in bins) > np.quantile(n, 0.98)
    assert sharding.make_bins([], 10) == []
    assert [b.tolist() for b in sharding.make_bins([50, 7], 10)] == [[0], [1]]     # a read larger than the budget is a bin of its own


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    sys.path.insert(0, ROOT)
    from dnascent_b200 import sharding as sh, synth as sy
    lengths = sy.lognormal_lengths(4000, 30_000, np.random.default_rng(11))      # same seed on every rank
    mine = sh.shard_reads(lengths, world)[rank]
    # what bench.py reduces: per-rank sample counts (SUM) and step times (MAX)
    load = torch.tensor([float(lengths[mine].sum()), float(mine.size)], dtype=torch.float64)
    tot = load.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    q.put((rank, load.tolist(), tot.tolist(), float(t.item()), gathered))
    dist.destroy_process_group()


def test_two_rank_shard_plan_gloo():
    """world_size 2 over gloo: ranks derive disjoint, complete, balanced shards from the same seed with no data-path
    collective; only the scalar reductions of bench.py (SUM of samples, MAX of time) cross ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, l0, tot0, t0, g0), (r1, l1, tot1, t1, g1) = out
    assert tot0 == tot1 and tot0[1] == 4000 and t0 == t1 == 2.0
    assert abs(l0[0] - l1[0]) / tot0[0] < 0.01
    assert g0 == g1 and sorted(g0[0] + g0[1]) == list(range(4000)) and not set(g0[0]) & set(g0[1])


def test_constant_divisor_sequences(tmp_path):
    """The tile kernel divides by the window length with q0 = x*r, rem = fma(-q0, w, x), q = fma(rem, r, q0)
    instead of the IEEE division subroutine (csrc/seg.cu, div_const).  Compiled here with hardware FMA and compared
    with `/`: every 16th float bit pattern with |x| >= 2^-124 and 2*10^6 doubles per divisor w = 2..7 (the exhaustive
    float run, stride 1, was done once: 0 mismatches)."""
    import subprocess
    exe = tmp_path / "constdiv_check"
    src = os.path.join(os.path.dirname(__file__), "helpers", "constdiv_check.c")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe), src, "-lm"])
    out = subprocess.check_output([str(exe), "16", "2000000"]).decode().split()
    assert out == ["0", "0"], out


def test_exact_rounded_prefix_sums_by_map_composition():
    """The design check behind DESIGN.md s.8 item 5 (tests/helpers/proto_exact_prefix.py): scrappie's sequentially rounded
    prefix sums (event_detection.c:35-48) reproduced bit for bit with one serial step per 512-sample tile instead of one
    per sample, on POD5-like and adversarial (ties, wide dynamic range) inputs."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "proto_exact_prefix.py")
    spec = importlib.util.spec_from_file_location("proto_exact_prefix", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(3)
    dac = rng.integers(200, 900, 60_000).astype(np.int16)
    mod.check("pod5-like", (dac.astype(np.float32) + np.float32(10.0)) * np.float32(0.1755))
    mod.check("ties", rng.integers(1, 64, 40_000) * 0.5)
    mod.check("wide", np.exp(rng.normal(0, 4, 40_000)))
    # the kernel-shaped version csrc/seg_scan.cu mirrors: 64-sample blocks, signed addends, maps under a predicted binade,
    # exact check + literal fallback
    mod.check_blocks("pod5-like", (dac.astype(np.float32) + np.float32(10.0)) * np.float32(0.1755))
    mod.check_blocks("signed with drift", rng.normal(3, 50, 30_000))
    mod.check_blocks("signed ties", rng.integers(-64, 64, 30_000) * 0.5)


def test_streaming_scan_model_matches_the_sequential_prefix_sums():
    """The algorithm of csrc/seg_scan.cu, modelled lane by lane on the CPU (tests/helpers/proto_seg_scan.py), against
    scrappie's sequential prefix sums on POD5-like and adversarial signals: checkpoints and totals bit-identical
    wherever the read-level exactness condition of `sum` holds (reads where it does not go to the serial kernel)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    import proto_seg_scan as ps
    rng = np.random.default_rng(7)

    def pod5_like(n, scale=0.1755, off=-240.0):
        lvl = np.repeat(rng.normal(650, 60, size=n // 8 + 1), 8)[:n]
        dac = np.rint(lvl + rng.normal(0, 6, size=n)).astype(np.int16)
        return ((dac.astype(np.float32) + np.float32(off)) * np.float32(scale)).astype(np.float32)

    cases = {
        "pod5_like": pod5_like(40_000),
        "short_tail": pod5_like(64 * 33 + 17),
        "one_block": pod5_like(40),
        "ties": (np.round(rng.uniform(8, 200, size=20_000) * 4) / 4).astype(np.float32),          # few mantissa bits: many exact ties
        "powers_of_two": np.ldexp(1.0, rng.integers(-3, 9, size=12_000)).astype(np.float32),
        "growing": (np.linspace(1, 3000, 15_000) * rng.uniform(0.9, 1.1, size=15_000)).astype(np.float32),   # crosses binades fast
        "spikes": np.where(rng.random(15_000) < 0.01, 30000.0, rng.normal(90, 10, size=15_000)).astype(np.float32),
        "zeros_and_negatives": np.where(rng.random(10_000) < 0.2, 0.0, rng.normal(0, 50, size=10_000)).astype(np.float32),
        "tiny_then_large": np.concatenate([np.full(3000, 2.0 ** -19, dtype=np.float32), pod5_like(6000)]),
    }
    total_fast = 0
    for name, x in cases.items():
        st = {}
        cs, cq, s, q, exact = ps.scan(x, st)
        ws, wq, s_ref, q_ref = ps.sequential(x)
        np.testing.assert_array_equal(cq, wq, err_msg=name)
        assert q == q_ref, name
        if exact:
            np.testing.assert_array_equal(cs, ws, err_msg=name)
            assert s == s_ref, name
        total_fast += st["fast_chunks"]
        if name == "pod5_like":
            assert exact and st["fast_chunks"] >= 0.6 * (st["fast_chunks"] + st["serial_chunks"]), st
    assert total_fast > 30
    # a signal whose sum cannot be exact in double (2^-20-sized samples next to 2^29-sized ones) must be flagged
    bad = np.concatenate([np.full(100, 2.0 ** -20 * 1.5, dtype=np.float32), np.full(100, 2.0 ** 29 * 1.25, dtype=np.float32)])
    assert not ps.scan(bad)[4]


def test_nan_sort_path_matches_std_sort(tmp_path):
    """dnascent_b200/csrc/nan_sort_path.cuh (where ONE NaN ends up under libstdc++'s std::sort, and the closed-form
    partition the device evaluates with prefix counts) against the real std::sort of this toolchain, the one the
    reference is built with: slope-like, duplicate-heavy, sorted and reversed inputs, 17 to 499 500 elements."""
    import shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "nan_sort_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-x", "c++", "-I", os.path.join(root, "dnascent_b200", "csrc"),
                    os.path.join(root, "oracle", "nan_sort_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "150"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok 150 trials"), out.stdout + out.stderr
    assert " not emulated 0;" in out.stdout and "150 checked" in out.stdout, out.stdout


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and prints ONE JSON line
    with the contract's keys: same metric / unit / config as our arm, impl == reference, a cpu_baseline describing the
    run, an e2e object with zero copy bytes, and no GPU launches."""
    import json, subprocess
    from oracle import refbind
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s through event-detect+banded-align" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == ("reference" if refbind.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"]
