#!/bin/bash
# N = 2: library multi-GPU test + the driver's torchrun line (short)
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --passes-per-step 20 --e2e-passes 12 --wieber-walks 296 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
tail -3 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
print('sweep', json.dumps(d['sweep'])[:1500])
print('wieber', d['wieber_front_to_back']['qp_periods_per_s'])
PY
