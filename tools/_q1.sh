timeout 600 python -m pytest tests/test_gpu_prepass.py tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/q1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q1_tests.log; tail -3 gpurun_out/q1_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/q1_bench.json 2> gpurun_out/q1_bench.err; echo bench rc=$?
python -c "
import json;d=json.loads(open('gpurun_out/q1_bench.json').read().strip().splitlines()[-1]);print(d['summary']);print(d['kernels_ms']);print(d['roofline']['kernel'],d['roofline']['frac'])"
