#!/bin/bash
# Round profile collection on the GPU box: GPU tests, the bench line, the launch list of the bench command, one --set full capture
# of each hot kernel, the non-headline configurations.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log)
timeout 300 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cd_lpc -c 2 -o gpurun_out/cd_lpc python tools/ncu_cd.py > gpurun_out/ncu_lpc.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cd_blk -c 1 -o gpurun_out/cd_blk python tools/blk_probe.py 200 512 20 4 grid > gpurun_out/ncu_blk.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:admm_res -c 1 -o gpurun_out/admm_res python tools/ncu_admm.py > gpurun_out/ncu_admm.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dgemm_mma -c 2 -o gpurun_out/dgemm_mma python tools/ncu_cd.py > gpurun_out/ncu_gemm.log 2>&1
timeout 300 python tools/configs_bench.py > gpurun_out/configs.log 2>&1
for a in "bls --n 40 --m 60 --samples 64" "maxcut --n 60 --p 0.15 --samples 64" "beam" "circle --n 5 --samples 32 --num-iters 30"; do timeout 60 python examples/suggest_and_improve.py $a 2>&1 | grep -v "^ \|^\[\|^var"; done > gpurun_out/examples.log 2>&1
for t in racecheck memcheck; do timeout 600 compute-sanitizer --tool $t python tools/sanitize_probe.py > gpurun_out/sanitizer_$t.log 2>&1; done
tail -3 gpurun_out/t_gpu.log; tail -1 gpurun_out/bench.log | cut -c1-400; tail -1 gpurun_out/bench_ref.log | cut -c1-600; tail -45 gpurun_out/configs.log
# With `gpurun --gpus N` (N > 1) additionally: the bench line and the facade's sharded batch mode under torchrun
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus N > gpurun_out/bench_nN.log 2>&1
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 examples/torchrun_batch.py --samples 2048 > gpurun_out/torchrun_batch_nN.log 2>&1
#   (the printed objective / restart index must equal the N = 1 run's)
