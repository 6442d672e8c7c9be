#!/bin/bash
# the driver's own two lines at N = 2: reference arm, then the CUDA arm, default flags
mkdir -p gpurun_out
s=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 > gpurun_out/bench_r2d_n2_reference.json 2> gpurun_out/bench_r2d_n2_reference.err
e=$(date +%s); echo "reference arm: $((e-s)) s, lines: $(wc -l < gpurun_out/bench_r2d_n2_reference.json)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/bench_r2d_n2.json 2> gpurun_out/bench_r2d_n2.err
f=$(date +%s); echo "cuda arm: $((f-e)) s, lines: $(wc -l < gpurun_out/bench_r2d_n2.json)"
tail -2 gpurun_out/bench_r2d_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d_n2.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_r2d_n2_reference.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'n_gpus', d['n_gpus'], r['n_gpus'])
print('herdt', d['herdt']['qp_solves_per_s'], 'sweep', d['sweep']['seconds'], d['sweep']['library_multi_gpu']['seconds'])
print('numa', d['e2e'].get('numa'), d['e2e']['host_link'])
PY
