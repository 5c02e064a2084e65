#!/bin/bash
# ncu evidence of round 2, second pass (bf16 configuration after the folded pre-activation, upsample2x and the
# back-to-back bottleneck kernel went in). Run on the GPU box through gpurun; summaries are copied to profiles/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_bf16_b128.csv python scripts/one_forward.py --precision bf16 --batch 128 --iters 3 > gpurun_out/ncu_a.log 2>&1
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed"
$NCU --metrics $M -k "regex:conv_tc_kernel|conv3x3_c64_halo_kernel|conv1x1_b2b_kernel" --csv --log-file gpurun_out/conv_ncu_r2_bf16.csv python scripts/one_forward.py --precision bf16 --batch 128 --iters 2 > gpurun_out/ncu_c.log 2>&1
for k in conv1x1_b2b_kernel upsample2x_kernel; do
  $NCU --set full --import-source on -k regex:$k -s 3 -c 3 -o gpurun_out/full_r2_$k -f python scripts/one_forward.py --precision bf16 --batch 128 --iters 2 > gpurun_out/ncu_full_$k.log 2>&1
done
# the PRE variant of conv_tc_kernel: its six launches of the second forward
$NCU --set full --import-source on -k "regex:conv_tc_kernel<128, 128, 2, 1>" -s 6 -c 6 -o gpurun_out/full_r2_conv_tc_pre -f python scripts/one_forward.py --precision bf16 --batch 128 --iters 2 > gpurun_out/ncu_full_pre.log 2>&1
for f in gpurun_out/full_r2_conv1x1_b2b_kernel.ncu-rep gpurun_out/full_r2_upsample2x_kernel.ncu-rep gpurun_out/full_r2_conv_tc_pre.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $f --page details > ${f%.ncu-rep}_details.txt 2>/dev/null
  if [ $(stat -c %s $f) -gt 9000000 ]; then rm -f $f; fi
done
ls -la gpurun_out/ | grep -E "r2_(conv1x1|upsample|conv_tc_pre|bf16)"; du -sh gpurun_out
