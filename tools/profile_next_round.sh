#!/bin/bash
# ncu launch lists for the next round (one GPU; never under torchrun).  Outputs under gpurun_out/, copy the summaries to profiles/.
#   gpurun --timeout 900 -- 'bash tools/profile_next_round.sh'
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
# (1) the bench command (kernel shares of a step)
$NCU -c 400 --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
# (2) batched vs single 512 x 512 SVD: where the time of one batched factorisation goes
python tools/bench_svd_batched.py 512 1,4,16,32 > gpurun_out/r02_svd_batched.jsonl 2> gpurun_out/r02_svd_batched.err
$NCU -s 2000 -c 1500 --log-file gpurun_out/r02_svd_batched_launches.csv python tools/bench_svd_batched.py 512 16 > /dev/null 2>&1
# (3) one H_eff kernel with full counters (DMMA pipe utilisation, DRAM bytes)
ncu --set full --import-source on --clock-control none -k regex:zgemm_sk_kernel -s 6 -c 2 -o gpurun_out/r02_zgemm_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
cat gpurun_out/r02_svd_batched.jsonl
