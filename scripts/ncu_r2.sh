#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/).
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare SHARES, not absolutes)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_bf16_b128.csv python scripts/one_forward.py --precision bf16 --batch 128 --iters 3 > gpurun_out/ncu_a.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_fp32_b32.csv python scripts/one_forward.py --precision fp32 --batch 32 --iters 3 > gpurun_out/ncu_b.log 2>&1
# per-launch tensor pipe / DRAM / L2->SM of every conv launch of the last forward (both configurations)
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed"
$NCU --metrics $M -k "regex:conv_tc_kernel|conv3x3_c64_halo_kernel|conv1x1_b2b_kernel" --csv --log-file gpurun_out/conv_ncu_r2_bf16.csv python scripts/one_forward.py --precision bf16 --batch 128 --iters 2 > gpurun_out/ncu_c.log 2>&1
$NCU --metrics $M -k regex:conv_tf32_kernel --csv --log-file gpurun_out/conv_ncu_r2_fp32.csv python scripts/one_forward.py --precision fp32 --batch 32 --iters 2 > gpurun_out/ncu_d.log 2>&1
# --set full of the joint-space tensor-core kernels (bf16 configuration, B=128) and of the fp32 attention conv
for k in gcn_gemm_tc_kernel bone_fusion_tc_kernel bone_coef_tc_kernel ste_tc_kernel regress_mano_kernel joint_embed_kernel conv3x3_c64_halo_kernel conv1x1_b2b_kernel; do
  $NCU --set full --import-source on -k regex:$k -s 2 -c 2 -o gpurun_out/full_r2_$k -f python scripts/one_forward.py --precision bf16 --batch 128 --iters 2 > gpurun_out/ncu_full_$k.log 2>&1
done
# the attention conv (3x3 2048->2048 @8x8, the largest launch) and its neighbours on the fp32 configuration's 3xTF32 kernel
$NCU --set full --import-source on -k regex:conv_tf32_kernel -s 142 -c 8 -o gpurun_out/full_r2_conv_tf32 -f python scripts/one_forward.py --precision fp32 --batch 32 --iters 2 > gpurun_out/ncu_full_tf32.log 2>&1
# text summaries (what gets committed under profiles/); gpurun merges at most 64 MiB back, so large reports are dropped
for f in gpurun_out/full_r2_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $f --page details > ${f%.ncu-rep}_details.txt 2>/dev/null
  if [ $(stat -c %s $f) -gt 9000000 ]; then rm -f $f; fi
done
ls -la gpurun_out/ | grep r2; du -sh gpurun_out
