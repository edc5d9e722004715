#!/bin/bash
# One gpurun call that re-measures everything profiles/ and BASELINE.md quote (run from the repo root
# on a B200 box):   gpurun --timeout 1500 -- 'bash tools/final_profile.sh <tag>'
# Afterwards, here:  python tools/summarize_profiles.py <tag>
# Numbers printed by the runs under ncu are never bench values; the bench lines come from the
# un-profiled runs at the top.
set -u
TAG=${1:-r2_final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python bench.py --steps 20 --warmup 5 > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"   # what the driver runs: config 2 + sub-records
python bench.py --config 2 --steps 20 --warmup 5 > "$OUT/bench_config2.json" 2> "$OUT/bench_config2.err"
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_config3.json" 2> "$OUT/bench_config3.err"
python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_config4.json" 2> "$OUT/bench_config4.err"
python bench.py --config 1 --steps 200 --warmup 20 --no-cpu-baseline > "$OUT/bench_config1.json" 2> "$OUT/bench_config1.err"
python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"
# launch list of the bench command (shares of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file "$OUT/launches_config2_2000scans.csv" \
  python bench.py --config 2 --scans 2000 --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
# one full-set capture of every kernel of one step
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_ring_runs|k_level_crop|k_surface_grid_cells|k_density|k_desc_hist|k_desc_mark|k_merge" -c 20 \
  -f -o "$OUT/prof_full" python bench.py --config 2 --scans 10000 --steps 1 --warmup 0 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
ncu -i "$OUT/prof_full.ncu-rep" --page raw --csv > "$OUT/ncu_full_raw_10000scans.csv" 2>/dev/null
python tools/parity_campaign.py > "$OUT/parity_campaign.log" 2>&1
python tools/parity_campaign.py 100000 > "$OUT/parity_campaign_shift.log" 2>&1
ls -la "$OUT"
