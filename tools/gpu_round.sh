#!/bin/bash
# One GPU-box session: parity tests, the bench line, the ncu launch list, one full capture of the hot kernel at the bench
# workload and one each at cfg 2 / cfg 3.   usage: tools/gpu_round.sh <tag> [pytest args]
tag=${1:-r02}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider "$@" > gpurun_out/tests_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?" >> gpurun_out/bench_$tag.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-strong"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu1_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcalf_fast_kernel -s 2 -c 1 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu2_$tag.log 2>&1
for cfg in 2 3; do
python tools/profile_cfg.py $cfg > gpurun_out/plain_cfg${cfg}_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mcalf_fast_kernel -s 2 -c 1 -f -o gpurun_out/prof_cfg${cfg}_$tag python tools/profile_cfg.py $cfg > gpurun_out/ncu_cfg${cfg}_$tag.log 2>&1
done
tail -3 gpurun_out/tests_$tag.log; cat gpurun_out/bench_$tag.json | head -c 600
