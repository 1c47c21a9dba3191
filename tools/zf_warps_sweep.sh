#!/bin/bash
# Launch-shape sweep of the Z / FIXED fast kernels (warps per block, -DLF_ZF_WARPS): build the variants in-tree first, e.g.
#   for w in 20 24; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
#       -DLF_ZF_WARPS=$w -o lumfuncmcmc_b200/csrc/liblfengine_w$w.so lumfuncmcmc_b200/csrc/lf_*.cu; done
# then run this script on a B200 (measured in round 2: 16 warps 1.086e12, 20 warps 1.055e12, 24 warps 0.994e12 terms/s at 1e6 x 512).
for lib in "" _w20 _w24; do
  for kind in z fixed; do
    if [ "$kind" = z ]; then ns=1e6; w=512; else ns=1e7; w=1024; fi
    for rep in 1 2; do
    LF_ENGINE_LIB=/root/repo/lumfuncmcmc_b200/csrc/liblfengine$lib.so timeout 300 python bench.py --kind $kind --nsources $ns --walkers $w --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('lib=$lib kind=$kind', d['value'], d['roofline']['frac'], d['ms_per_step'])
"
    done
  done
done
