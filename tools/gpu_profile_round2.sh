#!/bin/bash
# ncu evidence of round 2 (one B200): launch lists and --set full captures. Every profiled command runs once without ncu first.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-local --no-also --no-parity"
$B --workload c4 > gpurun_out/r02_plain_c4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_c4_2steps.csv $B --workload c4 > gpurun_out/r02_ncu_c4.log 2>&1
# the partitioned solver as a multi-GPU trial of config 5 runs it (12 fronts), one LM iteration after the lambda-init pass
P5="python tools/dev_gpu_profile_step.py 5 2 12"
$P5 > gpurun_out/r02_plain_c5_parts.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5_parts12_2iters.csv $P5 > gpurun_out/r02_ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel|pair_kernel|update_z_kernel|chol_band_kernel|panel_inverse|spike_forward|spike_gram|block_spike|block_gram' -s 3 -c 14 -o gpurun_out/r02_full_c5_parts12 $P5 > gpurun_out/r02_ncu_c5_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
