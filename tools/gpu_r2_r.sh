#!/bin/bash
for s in 1 2; do timeout 300 python tools/fuzz_host.py --seconds 120 --seed $s 2>&1 | tail -3; done
