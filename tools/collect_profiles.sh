#!/usr/bin/env bash
# Regenerate the round's measurement artifacts on a GPU box (run through gpurun); outputs land in gpurun_out/ and are summarised
# into profiles/ by tools/summarise_profiles.sh on the authoring side.
set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | head -c 300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>&1; tail -1 gpurun_out/bench_ref_final.json | head -c 200
python tools/conv_bench.py > gpurun_out/conv_bench_final.txt 2>&1
EGM_PROFILE_DETAIL=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph --profile-json gpurun_out/prof_final.json > /dev/null 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_pre.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
wc -l gpurun_out/launches_final.csv
