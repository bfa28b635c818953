#!/usr/bin/env python3
"""Multi-score (BASELINE.json configs[3]) timing on one B200: S score definitions over one resident
slab, tensor-core contraction (npc_multi.cuh) against the fused kernel run once per definition.
Wall clock around the synchronous npc_score_resident_multi call (row tables H2D, kernels, S x n
scores D2H all inside).  A genotype-score cell = (variant, sample, definition).

  python tools/bench_multi.py [--samples 200000] [--variants 697,20000] [--scores 18] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200_000)
    ap.add_argument("--variants", default="697,20000")
    ap.add_argument("--scores", type=int, default=18)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    import nimpress_b200 as nb
    from bench import SEED, cohort_params, make_rows

    dev = torch.device("cuda:0")
    n, S = args.samples, args.scores
    stride = -(-2 * n // 128) * 128
    for V in [int(v) for v in args.variants.split(",")]:
        af, beta, ref_is_ea, af_thr, miss_thr, alt = cohort_params(0, V)
        base = make_rows(nb.ROW_DTYPE, V, af, beta, ref_is_ea)
        rng = np.random.default_rng(7)
        lists = []
        for k in range(S):
            r = base.copy()
            r["beta"] = np.round(rng.normal(0, 0.05, V), 4)
            lists.append(r)
        offs = [0.0] * S
        eng = nb.Engine(n, max_rows_per_block=min(V, 32768), n_slots=0)
        gt = torch.empty((V, stride), dtype=torch.uint8, device=dev)
        eng.synth_fill_device(gt, stride, 0, V, SEED, torch.from_numpy(af_thr.view(np.int32)).to(dev),
                              torch.from_numpy(miss_thr.view(np.int32)).to(dev), torch.from_numpy(alt).to(dev))
        torch.cuda.synchronize()
        eng.resident_adopt(gt, stride, V)
        # outputs in pinned host memory, reused across calls: what a C caller holding its buffers sees
        pin_s = torch.empty((S, n), dtype=torch.float64).pin_memory()
        pin_l = torch.empty((S, V * nb.LOCUS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        sc_out = [pin_s[k].numpy() for k in range(S)]
        lo_out = [pin_l[k].numpy().view(nb.LOCUS_DTYPE) for k in range(S)]
        out = dict(samples=n, variants=V, scores=S, slab_gb=round(V * stride / 1e9, 3))
        res = {}
        for mode, env in (("contraction", "1"), ("one_by_one", "0")):
            os.environ["NPC_MULTI"] = env
            ts = []
            for rep in range(args.reps + 1):
                t0 = time.perf_counter()
                got = eng.score_resident_multi(lists, offs, sc_out, lo_out)
                ts.append(time.perf_counter() - t0)
            res[mode] = [(a.copy(), b, c.copy()) for a, b, c in got]
            t = float(np.median(ts[1:]))
            out[mode] = dict(ms=round(t * 1e3, 3), cells_per_s=V * n * S / t, genotype_bytes_per_s=2 * V * n / t)
        dev_rel = 0.0
        for k in range(S):
            a, b = res["contraction"][k][0], res["one_by_one"][k][0]
            assert res["contraction"][k][1] == res["one_by_one"][k][1]
            assert np.array_equal(res["contraction"][k][2], res["one_by_one"][k][2])
            dev_rel = max(dev_rel, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3))))
        out["speedup"] = round(out["one_by_one"]["ms"] / out["contraction"]["ms"], 2)
        out["max_rel_dev_between_paths"] = dev_rel
        out["contractions_served"] = eng.multi_contractions
        print(json.dumps(out))
        eng.close()
        del gt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
