timeout 600 python -m pytest tests/test_gpu_prepass.py -m gpu -x -q -k "bands_across" > gpurun_out/n2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/n2_tests.log; tail -5 gpurun_out/n2_tests.log
for M in peer nccl; do
HB_BANDS_EXCHANGE=$M timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_$M.json 2> gpurun_out/n2_bench_$M.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/n2_bench_$M.json').read().strip().splitlines()[-1]);print('$M', d['summary'])" || tail -5 gpurun_out/n2_bench_$M.err
done
