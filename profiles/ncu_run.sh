#!/bin/bash
# ncu --set full captures of the round's kernels (one GPU; each python run launches the kernel twice).
set -u
mkdir -p gpurun_out
cap() {  # name regex target
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 1 -c 1 -f -o gpurun_out/r02_$1 python profiles/ncu_targets.py $3 > gpurun_out/r02_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1.raw.csv 2>/dev/null
  python profiles/ncu_extract.py gpurun_out/r02_$1.raw.csv > gpurun_out/r02_$1.md 2>&1
}
cap wide_cfg5 fused_wide_kernel cfg5
cap wide_cfg3 fused_wide_kernel cfg3
cap fused_cfg2 fused_layer_kernel cfg2
cap inverse_cfg4 ar_inverse_kernel cfg4
cap rqs16 uni_kernel rqs16
ls -la gpurun_out | grep r02_
