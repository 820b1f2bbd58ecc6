timeout 900 python -m pytest tests/test_gpu_prepass.py -m gpu -x -q -k "pinned_planes" > gpurun_out/q1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q1_tests.log; tail -3 gpurun_out/q1_tests.log
for wl in 720p 1080p; do
timeout 600 python bench.py --mode intra_recon --workload $wl --steps 12 > gpurun_out/q1_ir_$wl.json 2> gpurun_out/q1_ir_$wl.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/q1_ir_$wl.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','identical_to_reference_reconstruction','gpu_launches')}, d['config']['workload'], d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('reference_whole_intra_encode_fps'))" || tail -5 gpurun_out/q1_ir_$wl.err
done
