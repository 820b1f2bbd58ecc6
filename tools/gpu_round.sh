#!/bin/bash
# One GPU lease: the -m gpu suite, the default bench line, the ncu launch list of one pre-pass frame and one --set full capture.
# usage (under gpurun): bash tools/gpu_round.sh TAG [tests|notests]
TAG=${1:-r02}; mkdir -p gpurun_out
if [ "${2:-tests}" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/profile_prepass.py 2 > gpurun_out/${TAG}_prof1.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_${TAG}.csv gpurun_out/inst_${TAG}.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_me|k_tq|k_subpel' -c 19 -f -o gpurun_out/prof_${TAG} python tools/profile_prepass.py 1 > gpurun_out/${TAG}_prof2.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
