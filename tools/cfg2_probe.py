#!/usr/bin/env python3
"""Why does bench.py's `extra.config2` (60 us) differ from tools/trace_config2.py (46 us)?  Same launch, timed both ways in one process."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nimpress_b200 as nb
import bench
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
for rep in range(2):
    t, out, shape, alg = bench.time_resident(nb, torch, dev, 100_000, 697, 0.005, 0.005)
    print("time_resident: %.1f us" % (t * 1e6), shape["grid"], shape["row_groups"])
    t, out, shape, alg = bench.time_resident(nb, torch, dev, 100_000, 697, 0.005, 0.005, steps=8, warmup=4)
    print("time_resident 8/4: %.1f us" % (t * 1e6))
