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
