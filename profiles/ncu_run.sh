#!/bin/bash
# ncu captures of the round's kernels (one GPU): `--set full` of one launch per kernel class (each python run
# launches the kernel twice, the second one is captured) and the launch list of the bench command.
set -u
mkdir -p gpurun_out
cap() {  # name regex target
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 1 -c 1 -f -o gpurun_out/r02_$1 python profiles/ncu_targets.py $3 > gpurun_out/r02_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1.raw.csv 2>/dev/null
  python profiles/ncu_extract.py gpurun_out/r02_$1.raw.csv > gpurun_out/r02_$1.md 2>&1
}
for t in "$@"; do
  case $t in
    cfg5) cap wide_cfg5 fused_wide_kernel cfg5;;
    cfg3) cap wide_cfg3 fused_wide_kernel cfg3;;
    cfg2) cap dual_cfg2 fused_dual_kernel cfg2;;
    cfg4) cap inverse_cfg4 ar_inverse_kernel cfg4;;
    rqs16) cap rqs16 uni_kernel rqs16;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.out 2>&1;;
  esac
done
ls -la gpurun_out | grep r02_ | head -40
