cd $GRAFT_REPO_ROOT
PATHWAYS=0,1,2,3,4,5,6,7,8,9,10,11 python tools/k50_probe.py 2>&1 | tail -12
ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sweep_launch_r2b.csv python tools/time_sweep.py > gpurun_out/ncu_r2b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/sweep_launch_r2b.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    if 'scb::' in r[ki]: print(r[ki][:60], r[vi])
PY
