#!/usr/bin/env python3
"""int16 GT storage through the fused pair kernel at the bench shard's sample width: kernel rate
against the format's own roofline -- 4 B per genotype (two int16 per diploid sample) -- with CUDA events.

    python tools/bench_int16.py [--samples 500000] [--variants 8192]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=500_000)
    ap.add_argument("--variants", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--width", type=int, default=2, choices=[2, 4], help="2: int16 (fused pair kernel); 4: int32 (generic two-kernel sequence)")
    args = ap.parse_args()
    import torch
    import nimpress_b200 as nb
    from bench import SEED, cohort_params, make_rows
    dev = torch.device("cuda:0")
    n, V = args.samples, args.variants
    stride8 = -(-2 * n // 128) * 128
    af, beta, ref_is_ea, af_thr, miss_thr, alt = cohort_params(0, V)
    rows = make_rows(nb.ROW_DTYPE, V, af, beta, ref_is_ea)
    e8 = nb.Engine(n, max_rows_per_block=V, n_slots=0)
    g8 = torch.empty((V, stride8), dtype=torch.uint8, device=dev)
    e8.synth_fill_device(g8, stride8, 0, V, SEED, torch.from_numpy(af_thr.view(np.int32)).to(dev),
                         torch.from_numpy(miss_thr.view(np.int32)).to(dev), torch.from_numpy(alt).to(dev))
    torch.cuda.synchronize()
    e8.close()
    W = args.width
    g16 = g8.to(torch.int16 if W == 2 else torch.int32)   # values 0..5: the same numbers, W bytes each
    del g8
    stride16 = g16.shape[1] * W
    eng = nb.Engine(n, ploidy=2, gt_width=W, max_rows_per_block=V, n_slots=0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_policy()
    d_rows = torch.from_numpy(rows.view(np.uint8).reshape(V, -1)).to(dev)
    ms = []
    for i in range(3 + args.steps):
        eng.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.score_block_device(g16, stride16, V, d_rows, n_rows=V)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms)) * 1e-3
    peak = 6547.2
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    alg = 2.0 * W * n * V + 32.0 * V + 16.0 * n
    print(json.dumps(dict(workload=f"{V} variants x {n} samples, int{8 * W} diploid GT ({2 * W} B/genotype), one call", launch_ms=t * 1e3,
                          genotypes_per_s=n * V / t, achieved_gbs=alg / t / 1e9, peak_gbs=peak, roofline_frac=alg / t / 1e9 / peak,
                          kernel_shape=eng.kernel_shape, nloci=eng.finish(want_loci=False)["nloci"])))
    eng.close()


if __name__ == "__main__":
    main()
