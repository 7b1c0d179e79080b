cd $GRAFT_REPO_ROOT
O=gpurun_out/r2k; mkdir -p $O
CMD="python bench.py --no-parity --no-e2e --no-cusparse --no-cpu-baseline --steps 3 --warmup 3 --iters-per-step 5 --shape 32,256,256"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 60 --csv --log-file $O/launches_slab8.csv $CMD > $O/ncu.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2k/launches_slab8.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[hi+1:hi+41]:
    print(r[kn][:70], r[mv])
PY
