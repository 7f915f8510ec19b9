#!/bin/bash
# Sharded DMRG sweeps on N GPUs (default 2): parity against the single-GPU sweep at chi=256, then a timed sweep on the C5 lattice.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/run_sharded_dmrg_2gpu.sh 2'
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
O=gpurun_out/r02_sharded_dmrg_$N
: > $O.jsonl; : > $O.err
timeout 300 $RUN tools/bench_multigpu.py --what dmrg --lx 6 --ly 4 --chi 256 --sweeps 2 --check >> $O.jsonl 2>> $O.err
timeout 300 $RUN tools/bench_multigpu.py --what dmrg --lx 6 --ly 4 --chi 256 --sweeps 2 --check --dist-svd >> $O.jsonl 2>> $O.err
timeout 600 $RUN tools/bench_multigpu.py --what dmrg --lx 12 --ly 6 --chi ${CHI:-1024} --sweeps 1 >> $O.jsonl 2>> $O.err
cat $O.jsonl; tail -5 $O.err
