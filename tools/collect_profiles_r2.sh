#!/bin/bash
# Round-2 profile collection on the GPU box (everything lands in gpurun_out/): launch list of the bench command, --set full captures
# of the hot kernels, the reference arm, racecheck text of the new kernel.
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cd_lpc2 -c 1 -f -o gpurun_out/cd_lpc2_r2 python tools/lpc2_probe.py 1000 1024 1 > gpurun_out/ncu_lpc2_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dgemm_mma -c 2 -f -o gpurun_out/dgemm_mma_r2 python tools/lpc2_probe.py 1000 1024 1 > gpurun_out/ncu_gemm_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cd_blk -c 1 -f -o gpurun_out/cd_blk_r2 python tools/blk_probe.py 200 512 20 4 grid > gpurun_out/ncu_blk_r2.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_probe.py lpc2 > gpurun_out/racecheck_lpc2_r2.log 2>&1
tail -c 1500 gpurun_out/bench_ref_r2.log; head -30 gpurun_out/racecheck_lpc2_r2.log | cut -c1-400
