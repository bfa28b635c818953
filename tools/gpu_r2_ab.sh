#!/bin/bash
# A/B on one box: the library of the previous commit (E) against the current one, the bench shard, steady state
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
cp nimpress_b200/lib/libnimpress_cuda.so /tmp/keep.so
for rep in 1 2; do
  for v in E cur; do
    if [ $v = E ]; then cp nimpress_b200/lib/variants/E.so nimpress_b200/lib/libnimpress_cuda.so; else cp /tmp/keep.so nimpress_b200/lib/libnimpress_cuda.so; fi
    timeout 600 $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'])"
  done
done
cp /tmp/keep.so nimpress_b200/lib/libnimpress_cuda.so
