timeout 600 python -m pytest tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/q1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q1_tests.log; tail -12 gpurun_out/q1_tests.log
