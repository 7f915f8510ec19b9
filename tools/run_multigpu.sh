#!/bin/bash
# Multi-GPU measurements of the two sharded paths on N GPUs of one box:
#   bash tools/run_multigpu.sh N      -> gpurun_out/multigpu_N.jsonl
# (1) MPO-bond-sharded H_eff matvec chi=2048, w=20 (and w=24 at N=8, where 20 does not divide evenly)
# (2) QJMC ensemble, C4 shapes (N=64, chi=256), 16 trajectories per GPU x 3 steps, 16 worker streams per GPU
N=${1:-2}
OUT=gpurun_out/multigpu_$N.jsonl
mkdir -p gpurun_out; : > $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 20 --steps 3 >> $OUT 2>> gpurun_out/multigpu_$N.err
if [ "$N" = "8" ]; then timeout 300 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 24 --steps 3 >> $OUT 2>> gpurun_out/multigpu_$N.err; fi
timeout 400 $RUN tools/bench_qjmc.py --sites 64 --chi 256 --traj $((16 * N)) --steps 3 --workers 16 >> $OUT 2>> gpurun_out/multigpu_$N.err
cat $OUT; tail -5 gpurun_out/multigpu_$N.err
