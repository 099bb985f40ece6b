#!/bin/bash
for d in 0 1 2 4 6 3 7; do echo "dbg=$d"; SPEEDY_K2_QUAD=1 SPEEDY_QDBG=$d timeout 120 python tools/bench_transforms.py 30 _dbg 2>&1 | grep grid_to_spec | tail -2 | cut -c1-110; done
