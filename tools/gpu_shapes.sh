#!/bin/bash
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --brief"
echo "config2 wood x 100k:";   timeout 200 $B --samples 100000 --variants 697 2>&1 | tail -1
echo "config4-ish 730 x 200k:"; timeout 200 $B --samples 200000 --variants 730 2>&1 | tail -1
echo "config5 10000 x 50k:";   timeout 200 $B --samples 50000 --variants 10000 2>&1 | tail -1
echo "1M samples x 4096:";     timeout 200 $B --samples 1000000 --variants 4096 --steps 5 2>&1 | tail -1
echo "2M samples x 2048 (two-kernel path):"; timeout 200 $B --samples 2000000 --variants 2048 --steps 5 2>&1 | tail -1
