#!/bin/bash
# Round-2 final pass on the GPU box after the cd_blk_kernel phase-1 rework (everything lands in gpurun_out/): GPU tests, the bench line
# (own and reference arm), launch list of the bench command, a --set full capture of cd_blk_kernel on C5's phase 1, the C5 line alone.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_gpu_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu_r2.log)
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_under_ncu_r2.log 2>&1
BLK_PROF=0 timeout 400 ncu --set full --import-source on --clock-control none -k regex:cd_blk -c 1 -f -o gpurun_out/cd_blk_p1_r2 python tools/blk_prof.py 512 8 > gpurun_out/ncu_blk_p1_r2.log 2>&1
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c5_r2.log 2> gpurun_out/bench_c5_r2.err
tail -3 gpurun_out/t_gpu_r2.log; tail -c 300 gpurun_out/bench_r2.err; tail -c 300 gpurun_out/bench_r2.log; tail -c 400 gpurun_out/bench_c5_r2.log
