cd /root/repo
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_prepass.py -m gpu -x -q -k "matches_oracle or subpel_planes" > gpurun_out/san_racecheck_b.log 2>&1; echo "racecheck b rc=$?"; tail -3 gpurun_out/san_racecheck_b.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_prepass.py -m gpu -x -q > gpurun_out/q1_tests.log 2>&1; tail -2 gpurun_out/q1_tests.log
for i in 1 2; do timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/q1_bench.json 2> gpurun_out/q1_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/q1_bench.json').read().strip().splitlines()[-1]);print(d['summary']['value_fps'], d['summary']['e2e_fps'], d['kernels_ms']['me'])"; done
