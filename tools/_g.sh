cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -1 gpurun_out/bench_r2j.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step']*1e3, 'e2e', d['e2e']['ms_per_step'], 'sweep', d['sweep']['seconds'], 'frac', d['roofline']['frac'], d['clocks'])
PY
