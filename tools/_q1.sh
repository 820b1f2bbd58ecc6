cd /root/repo
timeout 600 python -m pytest tests/test_gpu_prepass.py tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/q1_tests.log 2>&1; tail -2 gpurun_out/q1_tests.log
for i in 1 2; do timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/q1_bench.json 2> gpurun_out/q1_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/q1_bench.json').read().strip().splitlines()[-1]);print(d['summary']['value_fps'], d['summary']['e2e_fps'], d['summary']['one_stream_chain_fps'], d['kernels_ms']['me'])"; done
