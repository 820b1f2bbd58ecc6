timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q4_tests.log; tail -4 gpurun_out/q4_tests.log
timeout 900 python bench.py > gpurun_out/q4_bench.json 2> gpurun_out/q4_bench.err; echo bench rc=$?
python -c "
import json;d=json.loads(open('gpurun_out/q4_bench.json').read().strip().splitlines()[-1]);print(d['summary']);print(d['extra_workloads']);print(d['e2e']['value'])"
