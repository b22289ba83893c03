#!/bin/bash
# ncu evidence of round 2, final build (one B200): launch lists and --set full captures. Every profiled command runs once without ncu first.
# Under ncu the library sees the injection variables and queues the reduced-system solve AFTER the Schur pass (no overlap).
# (The launch lists of the whole bench command, r02f_launches_bench_c{5,4}_2steps.csv, were taken with the same script before the last two
# changes -- Dr records back to 128 bytes, 13 fronts -- and take 9 minutes of box time; this version profiles one resident solve instead.)
set -x
mkdir -p gpurun_out
for c in 5 4; do
  P="python tools/dev_gpu_profile_step.py $c 3 0"
  $P > gpurun_out/r02g_plain_c${c}_step.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_c${c}_3iters.csv $P > gpurun_out/r02g_ncu_c${c}.log 2>&1
done
# --set full: one LM trial after the lambda-init pass, every kernel of the trial once (config 5), the pass and update kernels of config 4
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel|tile_diag_kernel|pair_tile_mma_kernel|update_z_kernel|chol_band_kernel|spike_forward2|spike_gram|block_spike|block_gram' -s 2 -c 16 -o gpurun_out/r02g_full_c5 python tools/dev_gpu_profile_step.py 5 2 0 > gpurun_out/r02g_ncu_c5_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel|tile_diag_kernel|pair_tile_mma_kernel|update_z_kernel' -s 2 -c 4 -o gpurun_out/r02g_full_c4 python tools/dev_gpu_profile_step.py 4 2 0 > gpurun_out/r02g_ncu_c4_full.log 2>&1
ls -la gpurun_out/r02g*.ncu-rep
