#!/bin/bash
# Round-2 evidence run on a B200 box: GPU tests, the default bench line, the ncu launch list of the same command, and one
# `ncu --set full` capture each of the resident V_eff kernel (1e7 sources) and the z-evolving fast kernel (config 2 size).
# Every command is run plainly first; ncu follows only if that run exited 0 (profiles/README.md quotes the outputs).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_b1_final.json 2> gpurun_out/r2_b1_final.err; tail -c 300 gpurun_out/r2_b1_final.err
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_1e7x1024.csv \
      python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
python tools/veff_one.py 1e7 > gpurun_out/r2_vone_tma.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_veff_res_tma -s 1 -c 1 -o gpurun_out/r2_vres_tma_1e7 \
      python tools/veff_one.py 1e7 > /dev/null 2>&1
python bench.py --kind z --nsources 1e6 --walkers 512 --no-extras --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_zplain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_main -s 6 -c 1 -o gpurun_out/r2_kmain_z_1e6x512 \
      python bench.py --kind z --nsources 1e6 --walkers 512 --no-extras --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
