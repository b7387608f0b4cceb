cd $GRAFT_REPO_ROOT
ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sweep_launch_r2a.csv python tools/time_sweep.py > gpurun_out/ncu_r2a.log 2>&1
tail -2 gpurun_out/ncu_r2a.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/sweep_launch_r2a.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:]:
    print(r[ki][:70], r[vi], r[ui])
PY
