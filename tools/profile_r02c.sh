#!/bin/bash
# DRAM traffic / tensor-pipe activity of the two chi^3 stages of one H_eff application at the bench shape (a reduced metric list: the
# full set returned NaNs for these 164 ms launches)
R=/tmp/ncu_r02; mkdir -p $R gpurun_out
ncu --clock-control none --kernel-name-base demangled -k regex:'zgemm_kernel<.int.4, .int.1, .int.4, .int.4, .int.1>' -s 2 -c 2 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --csv --log-file gpurun_out/r02_matvec_stage_metrics.csv python tools/prof_matvec.py > /dev/null 2>&1
cat gpurun_out/r02_matvec_stage_metrics.csv | tail -14
