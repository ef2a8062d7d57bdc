#!/bin/bash
# Round-2 ncu evidence, run on the GPU box from the repo root (gpurun): launch list with labels, DRAM traffic of the GEMMs,
# and one `--set full` capture per hot kernel (reports are read back here with `ncu -i ... --page raw --csv`).
set -x
O=gpurun_out
SDB_LABEL_LOG=$O/r2f_labels.tsv ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file $O/r2f_launches.csv python profiles/profile_step.py 2 > $O/r2f_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tc --csv \
    --log-file $O/r2f_traffic.csv python profiles/profile_step.py 1 >> $O/r2f_ncu.log 2>&1
cap() {  # name, kernel regex (demangled), launches to skip. The report embeds the whole cubin (34 MB for the GEMM family) and
  # gpurun brings back at most 64 MiB: extract the raw metric page here, keep only the CSV
  ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o /tmp/r2f_$1 \
      python profiles/profile_step.py 2 >> $O/r2f_ncu.log 2>&1
  ncu -i /tmp/r2f_$1.ncu-rep --page raw --csv > $O/r2f_$1.raw.csv 2>> $O/r2f_ncu.log
  rm -f /tmp/r2f_$1.ncu-rep
}
cap attn48 "attention_kernel<.int.48, .int.2" 12
cap attn80 "attention_kernel<.int.80" 12
cap attn160 "attention_kernel<.int.160" 14
cap gn_apply "gn_apply_kernel" 70
cap conv_small3 "conv3x3_small_cout_kernel<.int.3>" 0
cap conv_small4 "conv3x3_small_cout_kernel<.int.4>" 1
cap conv_cin4 "conv3x3_cin4_kernel" 2
cap gemv "gemv_kernel" 4
cap gn_fold "gn_fold_kernel" 2
cap gemm_conv_l0 "gemm_tc_kernel<.int.160, .int.3, .int.3, .int.2, .int.1>" 25
cap gemm_geglu "gemm_tc_kernel<.int.128, .int.3, .int.2, .int.2, .int.3>" 11
cap gemm_vae512 "gemm_tc_kernel<.int.128, .int.1, .int.4, .int.2, .int.1>" 8
ls -la $O
