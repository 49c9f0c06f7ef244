#!/usr/bin/env bash
# Turn the raw files tools/collect_profiles.sh left in gpurun_out/ into the tracked summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
{ echo "# round 1 (final) -- ncu launch list of ONE EGM-UNet train step (batch 16, 480x480, bf16), eager launches of the same step bench.py times"; echo "# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2600 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph   (after the same command exited 0 without ncu)"; echo "# summarised by tools/ncu_step_summary.py; per-launch times are cold-cache and serialised: compare SHARES with the CUDA-event breakdown (step_breakdown_r1_events.txt)"; python tools/ncu_step_summary.py gpurun_out/launches_final.csv; } > profiles/launches_r1_summary.txt 2>&1
{ echo "# round 1 (final) -- CUDA-event breakdown of one eager train step per C-ABI entry point (bench.py --no-graph --profile-json, EGM_PROFILE_DETAIL=1; tools/prof_summary.py)"; python tools/prof_summary.py gpurun_out/prof_final.json 40; } > profiles/step_breakdown_r1_events.txt
{ echo "# round 1 (final) -- tools/conv_bench.py: tcgen05 conv kernels on the DoubleConv layer shapes, N=16, 3x3, bf16, CUDA events, L2 flushed between calls"; grep -v Warn gpurun_out/conv_bench_final.txt; } > profiles/conv_microbench_r1.txt
tail -1 gpurun_out/bench_final.json > profiles/bench_r1_1gpu.json
tail -1 gpurun_out/bench_ref_final.json > profiles/bench_r1_reference_arm.json
head -8 profiles/launches_r1_summary.txt | cut -c1-110
