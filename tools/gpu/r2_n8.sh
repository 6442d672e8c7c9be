#!/bin/bash
# N = 8 (gpurun --gpus 8): the driver's torchrun line restricted to the preview legs (device-resident value, end to end from host
# buffers with the ranks pinned to their GPU's NUMA node, host-link probe) and the configs[4] sweep (torchrun sharding and wg_multi)
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n${N}.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" > gpurun_out/lscpu_n${N}.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber > gpurun_out/bench_r2_n${N}.json 2> gpurun_out/bench_r2_n${N}.err
tail -3 gpurun_out/bench_r2_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_n${N}.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'pos', d['e2e']['com_position_only']['value'])
print('numa', d['e2e'].get('numa')); print('link', d['e2e'].get('host_link'))
print('sweep', json.dumps(d['sweep'])[:1500])
PY
