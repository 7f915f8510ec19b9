#!/bin/bash
# ncu --set full of the Jacobi-step kernels (tn_jacobi.cu), the pair EVD and the register-blocked Cholesky inside one N x N W-only factorisation
# (circle-method steps: TN_SVD_SPLIT=1, so a launch covers all pairs of a step), summarised by tools/ncu_summary.py into
# gpurun_out/r02d_ncu_jacobi_summary.json.  The skip counts step over the launches of the QR phase (2 QR steps x N/64 panels x 4 passes).
R=/tmp/ncu_r02d; mkdir -p $R gpurun_out
N=${1:-2048}
SK=$(( N / 64 * 8 + 44 ))
cap() {  # kernel regex, launches to skip, output name
  TN_SVD_SPLIT=1 timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base function -k regex:$1 -s $2 -c 1 -o $R/$3 -f \
      python tools/svd_once.py $N > $R/$3.log 2>&1
}
cap jacobi_gram64_kernel $SK gram64_step
cap jacobi_rot64_kernel $(( SK * 2 + 200 )) rot64_step
cap jacobi_rot64_kernel $(( N / 64 * 4 + 10 )) update64_qr
cap jacobi_evd64v2_kernel 10 evd64
cap chol_inv64c_kernel 10 chol64c
python tools/ncu_summary.py gpurun_out/r02d_ncu_jacobi_summary.json $R/*.ncu-rep > /dev/null 2>&1
