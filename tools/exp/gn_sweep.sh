#!/bin/bash
# GroupNorm streaming-kernel sweep (loads in flight x resident CTAs): one microbenchmark run per variant library
for f in tools/exp/libvar/libnhmc_*.so; do
  echo "== $f"
  NHMC_LIB_PATH=$PWD/$f timeout 200 python tools/kernel_microbench.py --only gn --resident 0 --solo 0 --reps 10 2>&1 | grep -E '"shape": \[8, (128|256), (256|128), ' 
done
