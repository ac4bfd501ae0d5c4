#!/bin/bash
# ncu --set full captures of the hot kernels (round 2).  usage: bash profiles/r2_capture.sh a|b
# (two calls: one gpurun call may bring back at most 64 MiB)
set -x
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --import-source on"
if [ "$1" = "d" ]; then
# very end of round: chain-only kernels after the lanes-per-row / compact-backward changes
timeout 300 python profiles/prof_step.py 64 256 8000 1 > gpurun_out/r2_plain_d64.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd_uni|bwd_uni)" -c 2 -o gpurun_out/r2_prof_d64_v3 python profiles/prof_step.py 64 256 8000 1 > gpurun_out/r2_ncu_d64.log 2>&1
timeout 300 python profiles/prof_step.py 128 64 8000 1 > gpurun_out/r2_plain_d128.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd_c4|bwd_c4)" -c 2 -o gpurun_out/r2_prof_d128_v3 python profiles/prof_step.py 128 64 8000 1 > gpurun_out/r2_ncu_d128.log 2>&1
elif [ "$1" = "c" ]; then
# end of round: D = 128 after the tensor-memory expectation kernel / two CTAs per SM forward, and the sampler
timeout 300 python profiles/prof_step.py 128 64 8000 1 > gpurun_out/r2_plain_d128.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd_c4|bwd_c4|sx2_tc|tiles_tc)" -c 6 -o gpurun_out/r2_prof_d128_v2 python profiles/prof_step.py 128 64 8000 1 > gpurun_out/r2_ncu_d128.log 2>&1
timeout 300 python profiles/prof_sampler.py 16000 > gpurun_out/r2_plain_sampler.log 2>&1 && timeout 600 $NCU -k regex:"psi_sample_kernel" -c 1 -o gpurun_out/r2_prof_sampler_v2 python profiles/prof_sampler.py 16000 > gpurun_out/r2_ncu_sampler.log 2>&1
elif [ "$1" = "a" ]; then
timeout 300 python profiles/prof_step.py 32 64 64000 1 > gpurun_out/r2_plain_cl.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd|bwd)_cl_kernel" -c 2 -o gpurun_out/r2_prof_cl_b python profiles/prof_step.py 32 64 64000 1 > gpurun_out/r2_ncu_cl.log 2>&1
timeout 300 python profiles/prof_step.py 64 148 16000 1 > gpurun_out/r2_plain_d64.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd_uni|bwd_uni|sx_tc|tiles_tc)" -c 5 -o gpurun_out/r2_prof_d64 python profiles/prof_step.py 64 148 16000 1 > gpurun_out/r2_ncu_d64.log 2>&1
else
timeout 300 python profiles/prof_step.py 128 37 8000 1 > gpurun_out/r2_plain_d128.log 2>&1 && timeout 600 $NCU -k regex:"psi_(fwd_c4|bwd_c4|sx2_tc|tiles_tc)" -c 6 -o gpurun_out/r2_prof_d128 python profiles/prof_step.py 128 37 8000 1 > gpurun_out/r2_ncu_d128.log 2>&1
timeout 300 python profiles/prof_sampler.py 16000 > gpurun_out/r2_plain_sampler.log 2>&1 && timeout 600 $NCU -k regex:"psi_sample_kernel" -c 1 -o gpurun_out/r2_prof_sampler python profiles/prof_sampler.py 16000 > gpurun_out/r2_ncu_sampler.log 2>&1
timeout 300 python profiles/prof_step.py 32 64 64000 1 2048 > gpurun_out/r2_plain_ck.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_ckpt_dram.csv python profiles/prof_step.py 32 64 64000 1 2048 > gpurun_out/r2_ncu_ck.log 2>&1
fi
