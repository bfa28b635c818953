"""The launch-shape choice of the tile kernel (npc_plan_shape: host arithmetic, no device) over the whole range of cohort
sizes.  Round 2's GPU fuzz found two holes in it after the fact -- a cohort whose best-scoring grid split did not fit
shared memory and silently fell back to the generic kernels, and 1.06-1.2 M samples for which no shape fitted at all
(0.28 of the roofline instead of 0.9); this test walks the sizes on the CPU instead."""
import numpy as np
import pytest

from nimpress_b200 import cuda

SMS, SMEM = 148, 232448                        # B200: SMs, cudaDevAttrMaxSharedMemoryPerBlockOptin
REGS = {16: 96, 24: 72, 28: 64}                # registers per thread of the kernel instances (__launch_bounds__ of npc_fused5.cuh)


def sizes():
    rng = np.random.default_rng(5)
    fixed = [1, 7, 8, 9, 255, 256, 257, 2000, 4096, 37_887, 37_888, 37_889, 157_929, 500_000, 1_060_864, 1_060_865, 1_212_416, 1_212_417]
    dense = list(range(1000, 1_400_000, 9973))
    return sorted(set(fixed + dense + [int(x) for x in rng.integers(1, 3_000_000, size=400)]))


def check(n, width, plan, exact, sms=SMS):
    mode = plan["mode"]
    if mode == 0:
        return
    K, nc, A, gd = plan["chunks_per_thread"], plan["consumer_warps"], plan["decider_warps"], plan["decider_tiles"]
    assert mode == (3 if plan["wide_slabs"] > 1 else 1 if exact else 2)
    assert K in (1, 2) and 1 <= nc <= plan["instance_warps"] and (K == 1 or nc <= 16), plan
    assert width == 1 or plan["instance_warps"] == 16, plan                       # int16 GT: the 16-warp instances only
    assert plan["threads"] == (nc + 2 + A) * 32 <= (plan["instance_warps"] + 4) * 32 <= 1024, plan
    assert plan["threads"] * REGS[plan["instance_warps"]] <= 65536, plan          # one CTA's registers fit the SM
    assert plan["sample_slabs"] * plan["row_groups"] <= sms, plan                 # cooperative launch: one CTA per SM
    assert plan["smem_bytes"] <= SMEM, plan
    assert gd in (4, 8) and plan["index_tiles"] >= gd + 2 and gd <= plan["lag"] <= plan["index_tiles"] - 1, plan
    assert plan["raw_stages"] >= (3 if K == 2 else 2), plan
    assert not exact or plan["row_groups"] == 1, plan
    # every chunk (8 samples) of a CTA's share of a row has a consumer lane
    n_slab = plan["wide_samples"] if mode == 3 else n
    chunks = -(-n_slab // 8)
    per_cta = -(-chunks // plan["sample_slabs"])
    assert per_cta <= 32 * K * nc and plan["slab_bytes"] == 32 * K * nc * 16 * width, plan
    if mode == 3:
        assert A == 1 and plan["wide_slabs"] * plan["wide_samples"] >= n and plan["wide_samples"] % 1024 == 0, plan


@pytest.mark.parametrize("width", [1, 2])
def test_every_cohort_size_gets_a_feasible_shape(width):
    lib_missing = None
    try:
        cuda.load_library()
    except cuda.NpcError as e:                  # the library is built by __graft_entry__.build(); the ABI test reports its absence
        lib_missing = e
    if lib_missing:
        pytest.skip(str(lib_missing))
    for n in sizes():
        for exact in (False, True):
            for n_rows in (64, 1 << 20):
                check(n, width, cuda.plan_shape(n, width, SMS, SMEM, n_rows, exact), exact)


@pytest.mark.parametrize("sms", [144, 132, 74, 16])
def test_other_sm_counts(sms):
    """A part with fewer SMs (another sm_100 SKU, a MIG slice): the same invariants, and still no cohort on the generic path."""
    for n in sizes()[::3]:
        for exact in (False, True):
            p = cuda.plan_shape(n, 1, sms, SMEM, 1 << 20, exact)
            assert p["mode"] != 0, (sms, n, p)
            check(n, 1, p, exact, sms)


def test_no_holes_in_the_fused_range():
    """int8 GT: every cohort up to 1,212,416 samples (148 CTAs x 16 warps x 2 chunks x 32 lanes x 8) runs the tile kernel
    in one resident pass, short and long launches alike; above that the decided-mode slabs, never the generic kernels."""
    for n in sizes():
        for n_rows in (64, 1 << 20):
            p = cuda.plan_shape(n, 1, SMS, SMEM, n_rows, False)
            assert p["mode"] == (2 if n <= 1_212_416 else 3), (n, p)
            e = cuda.plan_shape(n, 1, SMS, SMEM, n_rows, True)
            assert e["mode"] == (1 if n <= 1_212_416 else 3), (n, e)


def test_known_shapes():
    """The shapes the measurements in DESIGN 4.2b were taken on."""
    short = cuda.plan_shape(500_000, n_rows=697)
    long_ = cuda.plan_shape(500_000, n_rows=32768)
    assert (short["sample_slabs"], short["row_groups"], short["consumer_warps"]) == (148, 1, 14)
    assert (long_["sample_slabs"], long_["row_groups"], long_["consumer_warps"], long_["instance_warps"]) == (74, 2, 27, 28)
    p = cuda.plan_shape(900_000)
    assert (p["chunks_per_thread"], p["consumer_warps"], p["instance_warps"], p["decider_tiles"]) == (1, 24, 24, 4)
    p = cuda.plan_shape(1_180_000)
    assert (p["chunks_per_thread"], p["consumer_warps"], p["raw_stages"], p["decider_tiles"]) == (2, 16, 3, 4)
    p = cuda.plan_shape(100_000, n_rows=697)                                       # config 2: 29 slabs x 5 row groups
    assert (p["sample_slabs"], p["row_groups"], p["consumer_warps"]) == (29, 5, 14)
