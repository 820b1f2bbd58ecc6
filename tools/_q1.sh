for wl in 720p 1080p; do
timeout 600 python bench.py --mode intra_recon --workload $wl --steps 24 > gpurun_out/q1_ir_$wl.json 2> gpurun_out/q1_ir_$wl.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/q1_ir_$wl.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','identical_to_reference_reconstruction','gpu_launches')}, d['config']['workload'])" || tail -5 gpurun_out/q1_ir_$wl.err
done
timeout 300 python bench.py --mode intra_recon --workload 720p --steps 24 --streams 1 --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('1 stream', d['value'], d['ms_per_step'])"
