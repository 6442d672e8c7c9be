#!/bin/bash
# N=2: the driver's launch line for the scaling run + the reference arm under torchrun + the config-5 sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --sweep > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('herdt', d['herdt']['qp_solves_per_s'], d['herdt']['closed_loop_qp_solves_per_s'])
print('pldp', d['pldp']['pldp_solves_per_s']); print('dimitrov', d['dimitrov_front_to_back']['qp_periods_per_s'])
print('kajita', d['kajita_front_end']['preview_steps_per_s']); print('sweep', d['sweep'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_n8_reference.json 2>> gpurun_out/bench_n8.err; cut -c1-400 gpurun_out/bench_n8_reference.json
