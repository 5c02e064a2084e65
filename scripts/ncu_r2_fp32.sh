#!/bin/bash
# fp32-configuration part of scripts/ncu_r2.sh (re-run after conv_tf32.cu changed)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_fp32_b32.csv python scripts/one_forward.py --precision fp32 --batch 32 --iters 3 > gpurun_out/ncu_b.log 2>&1
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed"
$NCU --metrics $M -k regex:conv_tf32_kernel --csv --log-file gpurun_out/conv_ncu_r2_fp32.csv python scripts/one_forward.py --precision fp32 --batch 32 --iters 2 > gpurun_out/ncu_d.log 2>&1
# the attention conv is the 54th conv_tf32 launch of a forward (94 per forward): skip 94 + 50, capture 8
$NCU --set full --import-source on -k regex:conv_tf32_kernel -s 144 -c 8 -o gpurun_out/full_r2_conv_tf32 -f python scripts/one_forward.py --precision fp32 --batch 32 --iters 2 > gpurun_out/ncu_full_tf32.log 2>&1
ncu -i gpurun_out/full_r2_conv_tf32.ncu-rep --page raw --csv > gpurun_out/full_r2_conv_tf32_raw.csv 2>/dev/null
ncu -i gpurun_out/full_r2_conv_tf32.ncu-rep --page details > gpurun_out/full_r2_conv_tf32_details.txt 2>/dev/null
rm -f gpurun_out/full_r2_conv_tf32.ncu-rep
ls -la gpurun_out | grep -E "fp32|tf32"
