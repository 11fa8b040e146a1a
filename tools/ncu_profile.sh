#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one bench run + full-set captures of the top kernels,
# exported to CSV so that only text travels back. Usage: bash tools/ncu_profile.sh <round-tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_${TAG}.log 2>&1
for what in fprop wgrad dgrad; do
  ncu --set full --clock-control none --import-source on -k regex:"conv3d_march|wgrad_march" --launch-skip 2 -c 2 -o $OUT/prof_${TAG}_${what} \
      python tools/bench_layers.py $what 8 dec0b > $OUT/ncu_${what}_${TAG}.log 2>&1
  ncu -i $OUT/prof_${TAG}_${what}.ncu-rep --page raw --csv > $OUT/prof_${TAG}_${what}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${TAG}_${what}.ncu-rep --page details --csv > $OUT/prof_${TAG}_${what}_details.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:"maxpool3d_bwd|upsample3d_fwd|upsample3d_bwd|adam_kernel|bias_grad|first_wgrad|conv3d_first_kernel|head_bwd|head_fwd|maxpool3d_fwd|dice|repack_all" -c 48 -o $OUT/prof_${TAG}_bw \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_bw_${TAG}.log 2>&1
ncu -i $OUT/prof_${TAG}_bw.ncu-rep --page raw --csv > $OUT/prof_${TAG}_bw_raw.csv 2>/dev/null
ncu -i $OUT/prof_${TAG}_bw.ncu-rep --page details --csv > $OUT/prof_${TAG}_bw_details.csv 2>/dev/null
find $OUT -name "*.ncu-rep" -size +12M -delete
ls -la $OUT
