#!/bin/bash
# ncu evidence of round 2 (one B200): launch lists and --set full captures. Every profiled command runs once without ncu first.
# Under ncu the library sees the injection variables and queues the reduced-system solve AFTER the Schur pass (no overlap).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-local --no-also --no-parity"
# headline workload (config 5, one GPU: twelve fronts after the pass) and config 4 (two fronts): launch lists of the bench command
$B --workload c5 > gpurun_out/r02f_plain_c5.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_launches_bench_c5_2steps.csv $B --workload c5 > gpurun_out/r02f_ncu_c5.log 2>&1
$B --workload c4 > gpurun_out/r02f_plain_c4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_launches_bench_c4_2steps.csv $B --workload c4 > gpurun_out/r02f_ncu_c4.log 2>&1
# --set full: one LM trial after the lambda-init pass, every kernel of the trial once (config 5, then the pass kernels of config 4)
P5="python tools/dev_gpu_profile_step.py 5 2 0"
$P5 > gpurun_out/r02f_plain_c5_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel|tile_diag_kernel|pair_tile_mma_kernel|update_z_kernel|chol_band_kernel|spike_forward2|spike_gram|block_spike|block_gram' -s 2 -c 16 -o gpurun_out/r02f_full_c5 $P5 > gpurun_out/r02f_ncu_c5_full.log 2>&1
P4="python tools/dev_gpu_profile_step.py 4 2 0"
$P4 > gpurun_out/r02f_plain_c4_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel|tile_diag_kernel|pair_tile_mma_kernel|update_z_kernel' -s 2 -c 4 -o gpurun_out/r02f_full_c4 $P4 > gpurun_out/r02f_ncu_c4_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
