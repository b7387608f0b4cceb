cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i "numa\|socket" | head -6
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import torch
from scbonita2_b200 import device
print('affinity before', len(os.sched_getaffinity(0)))
print('bound', device.bind_host_to_device(), 'affinity after', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4])
PY
python tools/time_e2e.py
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sweep-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e', d['e2e']['ms_per_step'], d['e2e'].get('host_numa_bound_to_gpu'), 'dense', d['e2e_dense']['ms_per_step'], 'step', d['ms_per_step'], 'sweep e2e', d['sweep']['e2e']['seconds'])"
