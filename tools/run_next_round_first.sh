#!/bin/bash
# First GPU call of the next round (everything the last session of round 1 built without GPU time left):
#   gpurun --timeout 900 -- 'bash tools/run_next_round_first.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/run_next_round_first.sh 2' (sharded sweep at 2 GPUs)
# Writes gpurun_out/next_*.log / .jsonl.
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests -q -m gpu > gpurun_out/next_gpu_tests.log 2>&1; tail -5 gpurun_out/next_gpu_tests.log
  # the opt-in tests (device code that has never run on a GPU), one process each so that a crash in one cannot hide the others
  : > gpurun_out/next_unverified.log
  for T in qjmc_ensemble_batching_rounds sharded_heff_pipelined qjmc_front_end_with_observers applygates_fidelity itebd_step itebd_against_golden \
           thermal_example_energy applympo_matches tebd_with_projector three_call_svd balanced_sharded_heff; do
    echo "== $T" >> gpurun_out/next_unverified.log
    TN_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_z_projsum.py -q -k "$T" >> gpurun_out/next_unverified.log 2>&1
  done
  grep -E "^==|passed|failed|error" gpurun_out/next_unverified.log
  timeout 120 python tools/bench_multigpu.py --what dmrg --lx 6 --ly 4 --chi 256 --sweeps 2 --check > gpurun_out/next_sharded_dmrg_1.jsonl 2> gpurun_out/next_sharded_dmrg_1.err
  cat gpurun_out/next_sharded_dmrg_1.jsonl; tail -3 gpurun_out/next_sharded_dmrg_1.err
  # QJMC C4 shapes: plain multi-stream ensemble vs batching rounds
  timeout 300 python tools/bench_qjmc.py --sites 64 --chi 256 --traj 32 --steps 2 --workers 16 > gpurun_out/next_qjmc_plain.jsonl 2> gpurun_out/next_qjmc_plain.err
  TN_QJMC_BATCH=1 TN_QJMC_BATCH_STATS=1 timeout 300 python tools/bench_qjmc.py --sites 64 --chi 256 --traj 32 --steps 2 --workers 32 > gpurun_out/next_qjmc_batched.jsonl 2> gpurun_out/next_qjmc_batched.err
  cat gpurun_out/next_qjmc_plain.jsonl gpurun_out/next_qjmc_batched.jsonl; tail -2 gpurun_out/next_qjmc_batched.err
  timeout 200 python bench.py > gpurun_out/next_bench.json 2> gpurun_out/next_bench.err; cat gpurun_out/next_bench.json
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
  timeout 300 $RUN tools/bench_multigpu.py --what dmrg --lx 6 --ly 4 --chi 256 --sweeps 2 --check > gpurun_out/next_sharded_dmrg_$N.jsonl 2> gpurun_out/next_sharded_dmrg_$N.err
  timeout 300 $RUN tools/bench_multigpu.py --what dmrg --lx 6 --ly 4 --chi 256 --sweeps 2 --check --dist-svd >> gpurun_out/next_sharded_dmrg_$N.jsonl 2>> gpurun_out/next_sharded_dmrg_$N.err
  timeout 500 $RUN tools/bench_multigpu.py --what dmrg --lx 12 --ly 6 --chi 1024 --sweeps 1 >> gpurun_out/next_sharded_dmrg_$N.jsonl 2>> gpurun_out/next_sharded_dmrg_$N.err
  timeout 500 $RUN tools/bench_multigpu.py --what dmrg --lx 12 --ly 6 --chi 1024 --sweeps 1 --dist-svd >> gpurun_out/next_sharded_dmrg_$N.jsonl 2>> gpurun_out/next_sharded_dmrg_$N.err
  # sharded matvec: plain vs pipelined reduce_scatter (must print the same checksum)
  for P in 0 4 8; do timeout 200 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 24 --steps 3 --pipeline $P >> gpurun_out/next_sharded_heff_$N.jsonl 2>> gpurun_out/next_sharded_heff_$N.err; done
  # w = 20 (the C5 MPO): whole-bond-value split vs the balanced split
  timeout 200 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 20 --steps 3 >> gpurun_out/next_sharded_heff_$N.jsonl 2>> gpurun_out/next_sharded_heff_$N.err
  timeout 200 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 20 --steps 3 --balanced >> gpurun_out/next_sharded_heff_$N.jsonl 2>> gpurun_out/next_sharded_heff_$N.err
  timeout 200 $RUN tools/bench_multigpu.py --what heff --chi 2048 --w 20 --steps 3 --balanced --pipeline 4 >> gpurun_out/next_sharded_heff_$N.jsonl 2>> gpurun_out/next_sharded_heff_$N.err
  cat gpurun_out/next_sharded_dmrg_$N.jsonl gpurun_out/next_sharded_heff_$N.jsonl; tail -5 gpurun_out/next_sharded_dmrg_$N.err
fi
