"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: shard ranges, the
fixed-order combine of partial sums, the integer tally all-reduce and the slab gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nimpress_b200 import shard
    n, V = 1001, 97
    rng = np.random.default_rng(5)
    contrib = rng.normal(0, 0.05, size=(V, n))                      # row v's contribution to every sample
    lo, hi = shard.shard_range(V, world, rank)
    part = np.zeros(n)
    for v in range(lo, hi):                                          # this rank's left-to-right chain
        part += contrib[v]
    total, nl = shard.combine_partials(torch.from_numpy(part), torch.tensor([hi - lo], dtype=torch.int64))
    # sample-sharded pieces
    slo, shi = shard.shard_range(n, world, rank)
    counts = torch.from_numpy(np.stack([(contrib[:, slo:shi] > 0.05).sum(1), (contrib[:, slo:shi] < 0).sum(1)], 1).astype(np.int64))
    counts = shard.combine_counts(counts)
    sizes = [shard.shard_range(n, world, r)[1] - shard.shard_range(n, world, r)[0] for r in range(world)]
    gathered = shard.gather_scores(torch.from_numpy(part[slo:shi].copy()), sizes)
    q.put((rank, total.numpy(), int(nl.item()), counts.numpy(), gathered.numpy(), (lo, hi), part))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_combine_is_fixed_order_and_identical_on_all_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, V = 1001, 97
    rng = np.random.default_rng(5)
    contrib = rng.normal(0, 0.05, size=(V, n))
    # ranges tile [0, V) contiguously and are balanced
    ranges = [r[5] for r in res]
    assert ranges[0][0] == 0 and ranges[-1][1] == V and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    # combine = partials added in rank order, bit-identical on every rank
    want = res[0][6].copy()
    for r in res[1:]:
        want += r[6]
    for r in res:
        assert np.array_equal(r[1].view(np.uint64), want.view(np.uint64)) and r[2] == V
    # close to the single-chain sum (different association only)
    single = np.zeros(n)
    for v in range(V):
        single += contrib[v]
    assert np.max(np.abs(want - single)) < 1e-13
    # integer tallies: exact; slabs: concatenation in rank order
    wc = np.stack([(contrib > 0.05).sum(1), (contrib < 0).sum(1)], 1)
    for r in res:
        assert np.array_equal(r[3], wc)
    slabs = []
    from nimpress_b200.shard import shard_range
    for k, r in enumerate(res):
        lo, hi = shard_range(n, world, k)
        slabs.append(r[6][lo:hi])
    for r in res:
        assert np.array_equal(r[4], np.concatenate(slabs))


def test_shard_range_edges():
    from nimpress_b200.shard import shard_range
    assert [shard_range(5, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    assert shard_range(0, 2, 1) == (0, 0) and shard_range(10, 1, 0) == (0, 10)
    cover = [shard_range(1_000_000, 8, r) for r in range(8)]
    assert cover[0] == (0, 125000) and cover[-1] == (875000, 1000000)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a reduced cohort:
    one JSON line with the contract's keys, impl = reference, no GPU involved."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--samples", "20000", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    j = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "workload" in j["config"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
