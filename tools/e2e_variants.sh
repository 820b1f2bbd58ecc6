run() { echo "== $1"; env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'threads', d['e2e']['host_threads'])"; }
nproc; lscpu | grep -E "NUMA|Socket|Thread|^CPU\(s\)"; nvidia-smi topo -m | head -14
run "HB_E2E_THREADS=4" 29521
run "HB_E2E_THREADS=4 HB_PIN_CPUS=1" 29522
run "HB_E2E_THREADS=8 HB_PIN_CPUS=1" 29523
run "HB_E2E_THREADS=8 HB_BLOCKING=1" 29524
run "HB_E2E_THREADS=2 HB_PIN_CPUS=1" 29525
