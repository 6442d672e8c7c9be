mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_r2_n4.json)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n4.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'n', d['n_gpus'], 'herdt', d['herdt']['qp_solves_per_s'], 'sweep', d['sweep']['seconds'], d['sweep']['library_multi_gpu']['seconds'], 'launches', d['gpu_launches'])
PY
