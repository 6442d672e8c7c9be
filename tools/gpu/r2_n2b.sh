#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --passes-per-step 4 --e2e-passes 4 --no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber > gpurun_out/bench_r2_n2b.json 2> gpurun_out/bench_r2_n2b.err
tail -3 gpurun_out/bench_r2_n2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2b.json').read().strip().splitlines()[-1])
print('sweep', d['sweep']['seconds'], json.dumps(d['sweep']['library_multi_gpu']))
PY
