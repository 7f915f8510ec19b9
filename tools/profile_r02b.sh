#!/bin/bash
# ncu --set full of the GEMM-shaped kernels (dominant contraction kernel at the bench shape, K = 64 rotation and Gram GEMMs of the SVD);
# the reports stay on the box (gpurun copies back at most 64 MiB), only the JSON summary returns.
NCU="ncu --clock-control none"
R=/tmp/ncu_r02; mkdir -p $R gpurun_out
$NCU --set full --kernel-name-base demangled -k regex:'zgemm_kernel<.int.4, .int.1, .int.4, .int.4, .int.1>' -s 1 -c 2 -o $R/matvec_main python tools/prof_matvec.py > /dev/null 2>&1
$NCU --set full --kernel-name-base demangled -k regex:'zgemm_kernel<.int.4, .int.2, .int.4, .int.4, .int.2>' -s 40 -c 1 -o $R/rot python tools/bench_svd.py 2048 graded > /dev/null 2>&1
$NCU --set full --kernel-name-base demangled -k regex:'zgemm_kernel<.int.2, .int.4, .int.4, .int.2, .int.1>' -s 200 -c 1 -o $R/gram python tools/bench_svd.py 2048 graded > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_ncu_gemm_summary.json $R/matvec_main.ncu-rep $R/rot.ncu-rep $R/gram.ncu-rep > /dev/null 2>&1
ls -la $R gpurun_out/r02_ncu_gemm_summary.json
