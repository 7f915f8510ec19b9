#!/bin/bash
# ncu --set full of the Jacobi-step kernels (tn_jacobi.cu) and the register-blocked Cholesky inside one 4096^2 / 2048^2 W-only factorisation,
# summarised by tools/ncu_summary.py into gpurun_out/r02d_ncu_jacobi_summary.json; plus the phase split at 4096.
R=/tmp/ncu_r02d; mkdir -p $R gpurun_out
N=${1:-4096}
for k in jacobi_gram64_kernel jacobi_rot64_kernel jacobi_evd64v2_kernel chol_inv64c_kernel; do
  TN_SVD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base function -k regex:$k -s 40 -c 1 -o $R/$k -f \
      python tools/svd_once.py $N > $R/$k.log 2>&1
done
python tools/ncu_summary.py gpurun_out/r02d_ncu_jacobi_summary.json $R/*.ncu-rep > /dev/null 2>&1
cp $R/jacobi_gram64_kernel.ncu-rep gpurun_out/r02d_gram64.ncu-rep 2>/dev/null
TN_SVD_PROFILE=1 timeout 200 python tools/bench_svd.py $N graded 2>&1 | grep svd_profile | awk "NR==2"
