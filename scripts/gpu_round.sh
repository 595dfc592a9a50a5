#!/bin/bash
# One gpurun call that produces everything profiles/ holds for a round (run from the repository root ON THE GPU BOX):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r02'
# Order follows the profiling recipe: each ncu capture only after its own command has exited 0 without ncu; numbers
# printed under ncu are never bench values.  Outputs land in gpurun_out/<tag>_*; copy what is to be judged to profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export SITATOR_PROGRESSBAR=false

echo "== 1. GPU test suite"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/${TAG}_pytest_gpu.log

echo "== 2. bench (plain, no profiler)"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err
BRC=$?; echo "bench rc=$BRC"; cut -c1-400 $OUT/${TAG}_bench_plain.json
[ $BRC -eq 0 ] || exit 1

echo "== 3. ncu launch list of the same command (kernel SHARES of the step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-frames 0 > $OUT/${TAG}_ncu_list.log 2>&1
python scripts/launch_list_md.py $OUT/${TAG}_launches.csv "bench.py launch list ($TAG)" > $OUT/${TAG}_launch_list_bench.md || true

echo "== 4. ncu --set full of the fused fill+assign kernel (first tier, one launch)"
timeout 300 python scripts/time_two_tier.py llzo 20000 2 > $OUT/${TAG}_tt_plain.json 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assign_fast -s 2 -c 1 \
    -o $OUT/${TAG}_k1_full -f python scripts/time_two_tier.py llzo 20000 2 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_k1_full.ncu-rep --page raw --csv > $OUT/${TAG}_k1_ncu_full.csv 2>/dev/null || true

echo "== 5. other shapes + dotprod report"
timeout 600 python scripts/config_sweep.py > $OUT/${TAG}_config_sweep.log 2> $OUT/${TAG}_config_sweep.err || true
cp $OUT/config_sweep.json $OUT/${TAG}_config_sweep.json 2>/dev/null || true      # (stdout is one JSON line per shape; the file is the list)
timeout 300 python scripts/dotprod_report.py > $OUT/${TAG}_dotprod.json 2> $OUT/${TAG}_dotprod.err || true
timeout 120 python scripts/e2e_phases.py 100000 $OUT/${TAG}_e2e_1gpu_phases.json > $OUT/${TAG}_e2e_1gpu_phases.txt 2>&1 || true
ls -la $OUT | tail -n 20
