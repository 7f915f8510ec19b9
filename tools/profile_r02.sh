#!/bin/bash
# ncu captures for round 2 (one GPU; never under torchrun).  Outputs under gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into
# profiles/r02_ncu_full_summary.json.
#   gpurun --timeout 1200 -- 'bash tools/profile_r02.sh'
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of two resident H_eff applications at the bench shape (kernel shares of a step)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_matvec_launches.csv python tools/prof_matvec.py > /dev/null 2>&1
# (2) the dominant contraction kernel at the bench shape (chi = 2048, w = 20): the two chi^3 stages of one application
$NCU --set full --import-source on --kernel-name-base demangled -k regex:'zgemm_kernel<.int.4, .int.1, .int.4, .int.4, .int.1>' -s 1 -c 2 -o gpurun_out/r02_matvec_main_full python tools/prof_matvec.py > /dev/null 2>&1
# (3) the SVD kernels at n = 2048: pair EVD, K = 64 rotation GEMM, Gram GEMM, blocked panel Cholesky
$NCU --set full --import-source on -k regex:jacobi_evd64v2 -s 40 -c 1 -o gpurun_out/r02_evd_full python tools/bench_svd.py 2048 graded > /dev/null 2>&1
$NCU --set full --import-source on --kernel-name-base demangled -k regex:'zgemm_kernel<.int.4, .int.2, .int.4, .int.4, .int.2>' -s 40 -c 1 -o gpurun_out/r02_rot_full python tools/bench_svd.py 2048 graded > /dev/null 2>&1
$NCU --set full --import-source on --kernel-name-base demangled -k regex:'zgemm_kernel<.int.2, .int.4, .int.4, .int.2, .int.1>' -s 200 -c 1 -o gpurun_out/r02_gram_full python tools/bench_svd.py 2048 graded > /dev/null 2>&1
$NCU --set full --import-source on -k regex:chol_inv64b -s 40 -c 1 -o gpurun_out/r02_chol_full python tools/bench_svd.py 2048 graded > /dev/null 2>&1
# (4) the single-CTA small-bond kernels inside a C1 sweep
$NCU --set full --import-source on -k regex:'lanczos_small|small_svd' -s 300 -c 2 -o gpurun_out/r02_small_full python tools/bench_dmrg.py c1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
