#!/usr/bin/env bash
# Round 2: regenerate the measurement artifacts on a GPU box (run through gpurun; raw files land in gpurun_out/), then -- on the authoring
# side, where gpurun_out/ has been merged back -- run this script with `summarise` to refresh the tracked summaries under profiles/.
#   gpurun --timeout 1500 -- tools/collect_profiles_r2.sh          (1 GPU, ~6 min)
#   tools/collect_profiles_r2.sh summarise
set -x
cd "$(dirname "$0")/.."
if [[ "${1:-}" == "summarise" ]]; then
  grep "^{" gpurun_out/bench_r2_1gpu.json > profiles/bench_r2_1gpu.json
  grep "^{" gpurun_out/bench_r2_reference_arm.json > profiles/bench_r2_reference_arm.json
  python tools/prof_summary.py gpurun_out/r2_final_prof.json 30 | head -34 > profiles/step_breakdown_r2_events.txt
  python tools/ncu_conv_traffic.py gpurun_out/launches_r2.csv gpurun_out/r2_calls_ncu.txt profiles/conv_dram_traffic_r2.json profiles/launches_r2_summary.txt
  cp gpurun_out/configs_r2.txt profiles/configs_r2.txt
  exit 0
fi
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_final_tests.log 2>&1; tail -2 gpurun_out/r2_final_tests.log
python bench.py --profile-json gpurun_out/r2_final_prof.json > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2>&1
python tools/run_configs.py > gpurun_out/configs_r2.txt 2>&1
# launch list of ONE step with DRAM bytes (only after the plain run exited 0), plus the call log that names each conv launch's layer
python tools/one_step.py > gpurun_out/one_step_plain.log 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_r2.csv python tools/one_step.py --call-log gpurun_out/r2_calls_ncu.txt > gpurun_out/ncu_launches.log 2>&1
# ncu --set full of the conv kernels (summarised by hand into profiles/conv_kernels_r2_ncu.txt)
python tools/conv_bench.py --shapes 512,256,60 32,32,480 --iters 2 > gpurun_out/convb_plain.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:'k_conv_tc|k_wgrad_tc' -c 12 -o gpurun_out/prof_conv_r2 \
      python tools/conv_bench.py --shapes 512,256,60 32,32,480 --iters 2 > gpurun_out/ncu_conv_r2.log 2>&1
