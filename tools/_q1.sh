timeout 900 python -m pytest tests/test_gpu_intra.py -m gpu -x -q > gpurun_out/q1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q1_tests.log; tail -6 gpurun_out/q1_tests.log
