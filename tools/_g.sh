cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2j_n2.json 2> gpurun_out/bench_r2j_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j_n2.json').read().strip().splitlines()[-1])
print('N=2 step us', d['ms_per_step']*1e3, 'sweep', d['sweep']['seconds'], 'strong', {k:d['strong'][k] for k in ('ga_ms_per_step','sweep_seconds')}, 'e2e', d['e2e']['ms_per_step'])
PY
