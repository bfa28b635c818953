"""Worker of test_nccl_combine_two_ranks (launched by torchrun, one rank per GPU): every rank scores its
contiguous range of the score rows, npc_comm_combine adds the partial sums in rank order over NCCL; every rank
checks the result bit for bit against the rank-order sum of the ranges scored alone on its own GPU, and
against the oracle's single chain within the re-association tolerance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import nimpress_b200 as nb
    import orc
    from util_cohort import assert_loci_equal, bits, random_cohort, random_rows, score_excess

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # only to hand the NCCL id around: the data path is npc_comm_*
    rng = np.random.default_rng(4242)                     # same cohort on every rank
    n, V = 7001, 160
    gt = random_cohort(rng, n, V, miss_rate=0.02)
    rows = random_rows(rng, V, n_rows=V + 11)
    cuts = [len(rows) * k // world for k in range(world + 1)]
    ident = [nb.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    for exact in (True, False):
        eng = nb.Engine(n, max_rows_per_block=512, n_slots=2, device=local)
        eng.set_policy(); eng.set_exact_order(exact); eng.reset()
        eng.comm_init(ident[0], rank, world)
        eng.score_host(gt, rows[cuts[rank]:cuts[rank + 1]])
        scores, nloci = eng.comm_combine(offset=-0.25)
        raw, nloci2 = eng.comm_combine(offset=None)
        # the same ranges alone on this GPU
        total, nl = None, 0
        for k in range(world):
            e = nb.Engine(n, max_rows_per_block=512, n_slots=2, device=local)
            e.set_policy(); e.set_exact_order(exact); e.reset()
            e.score_host(gt, rows[cuts[k]:cuts[k + 1]])
            p = e.partial()
            total = p["sums"] if total is None else total + p["sums"]
            nl += p["nloci"]
            e.close()
        ok = ~np.isnan(total)
        assert nloci == nloci2 == nl, (nloci, nloci2, nl)
        assert np.array_equal(np.isnan(raw), ~ok) and np.array_equal(bits(raw[ok]), bits(total[ok])), "combined sums differ from the rank-order sum"
        assert np.array_equal(bits(scores[ok]), bits(eng.normalise(total, nl, -0.25)[ok]))
        want = orc.score_matrix(gt, n, 2, rows.astype(orc.ROW_DTYPE), offset=-0.25)
        assert nloci == want["nloci"] and score_excess(scores, want["scores"], want) <= 1.0
        got = [None] * world
        dist.all_gather_object(got, bits(scores).tobytes())
        assert all(g == got[0] for g in got), "ranks disagree on the combined scores"
        eng.close()
        # NCCL ids are single use: a fresh one per communicator
        ident = [nb.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
    # ---- sample-sharded: every rank holds a slab of the samples for every row; the integer tallies are summed over
    # NCCL (npc_comm_sum_counts) before the decision, nothing floating-point crosses ranks, scores are concatenated
    s_lo, s_hi = n * rank // world // 8 * 8, (n * (rank + 1) // world // 8 * 8 if rank + 1 < world else n)
    sub = np.ascontiguousarray(gt[:, 2 * s_lo:2 * s_hi])
    pad = (-sub.shape[1]) % 16
    sub = np.pad(sub, ((0, 0), (0, pad)))
    ident = [nb.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    eng = nb.Engine(s_hi - s_lo, max_rows_per_block=512, n_slots=0, device=local)
    eng.set_policy(); eng.reset(); eng.comm_init(ident[0], rank, world); eng.set_cohort_size(n)
    d_gt = torch.from_numpy(sub.view(np.uint8)).cuda()
    counts = torch.zeros((len(rows), 2), dtype=torch.int64, device="cuda")
    eng.count_block_device(d_gt, d_gt.shape[1], V, rows, counts)
    eng.comm_sum_counts(counts, len(rows))
    eng.accumulate_block_device(d_gt, d_gt.shape[1], V, rows, counts)
    got = eng.finish(offset=-0.25)
    want = orc.score_matrix(gt, n, 2, rows.astype(orc.ROW_DTYPE), offset=-0.25)
    assert got["nloci"] == want["nloci"]
    assert_loci_equal(got["loci"], want["loci"])
    a, b = got["scores"], want["scores"][s_lo:s_hi]
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(bits(a[~np.isnan(a)]), bits(b[~np.isnan(b)])), "sample-sharded scores differ"
    eng.close()
    dist.barrier()
    if rank == 0:
        print("nccl combine ok: %d ranks, %d samples, %d rows" % (world, n, len(rows)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
