#!/bin/bash
# Round-2 profile collection on the GPU box (everything lands in gpurun_out/): GPU tests, the bench line (own and reference arm),
# launch list of the bench command, --set full captures of the hot kernels, sanitizer passes over every kernel family.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_gpu_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu_r2.log)
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cd_lpc2 -c 1 -f -o gpurun_out/cd_lpc2_r2 python tools/lpc2_probe.py 1000 1024 1 > gpurun_out/ncu_lpc2_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dgemm_mma -c 2 -f -o gpurun_out/dgemm_mma_r2 python tools/lpc2_probe.py 1000 1024 1 > gpurun_out/ncu_gemm_r2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"cd_lpc_kernel|lpc_p1_pre" -c 2 -f -o gpurun_out/cd_lpc_p1_r2 python tools/lpc2_probe.py 1000 1024 1 > gpurun_out/ncu_p1_r2.log 2>&1
for t in racecheck memcheck; do for w in lpc2 pipe cd blk admm; do echo "== $t $w"; timeout 500 compute-sanitizer --tool $t python tools/sanitize_probe.py $w 2>&1 | grep -v "^$" | cut -c1-400 | tail -12; done; done > gpurun_out/sanitizer_r2.log 2>&1
tail -3 gpurun_out/t_gpu_r2.log; tail -c 400 gpurun_out/bench_r2.err; tail -c 300 gpurun_out/bench_r2.log; grep "SUMMARY" gpurun_out/sanitizer_r2.log
