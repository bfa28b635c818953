#!/usr/bin/env python3
"""Where a short launch spends its time (BASELINE configs[1]: 697 loci x 100,000 samples, one launch): CUDA-event
time of the launch, and the %globaltimer stamps of CTA 0 inside it (NPC_TRACE=1)."""
import json
import os
import sys

import numpy as np

os.environ["NPC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import nimpress_b200 as nb
    from bench import SEED, cohort_params, make_rows
    dev = torch.device("cuda:0")
    out = []
    for n, V in ((100_000, 697), (200_000, 730), (50_000, 10_000), (500_000, 4096)):
        stride = -(-2 * n // 128) * 128
        af, beta, ref_is_ea, af_thr, miss_thr, alt = cohort_params(0, V)
        rows = make_rows(nb.ROW_DTYPE, V, af, beta, ref_is_ea)
        eng = nb.Engine(n, max_rows_per_block=V, n_slots=0)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        eng.set_stream(stream.cuda_stream)
        eng.set_policy()
        gt = torch.empty((V, stride), dtype=torch.uint8, device=dev)
        eng.synth_fill_device(gt, stride, 0, V, SEED, torch.from_numpy(af_thr.view(np.int32)).to(dev),
                              torch.from_numpy(miss_thr.view(np.int32)).to(dev), torch.from_numpy(alt).to(dev))
        d_rows = torch.from_numpy(rows.view(np.uint8).reshape(V, -1)).to(dev)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        ms, tr = [], []
        for i in range(12):
            eng.reset()
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.score_block_device(gt, stride, V, d_rows, n_rows=V)
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 4:
                ms.append(e0.elapsed_time(e1))
                t = eng.trace()
                tr.append([(x - t[0]) / 1e3 for x in t[1:6]])
        tr = np.median(np.array(tr), axis=0)
        alg = 2.0 * n * V + 32.0 * V + 16.0 * n
        out.append(dict(samples=n, loci=V, shape=eng.kernel_shape, launch_us=float(np.median(ms)) * 1e3, hbm_floor_us=alg / 6547.2e9 * 1e6,
                        cta0_us_after_start=dict(tables_ready=tr[0], first_tile_counted=tr[1], last_tile_counted=tr[2], last_tile_accumulated=tr[3], sums_stored=tr[4])))
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
