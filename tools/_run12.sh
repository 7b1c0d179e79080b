cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
for n in 2097152 4194304 8388608 16777216; do timeout 120 python tools/probe_vec.py $n; done 2>&1 | tee $O/vec.log
B="timeout 300 python bench.py --no-parity --no-e2e --no-cusparse --no-cpu-baseline --steps 20 --shape 32,256,256"
for mb in 0 20 40 60 90; do
  LSK_WS_RESIDENT_MB=$mb $B > $O/res_$mb.log 2>$O/res_$mb.err
  python - $O/res_$mb.log $mb <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("resident MB", sys.argv[2], round(d["value"],1), "it/s", round(1e6/d["value"],2), "us/it", "spmv alone", round(d["roofline"]["ms_per_launch"]*1e3,2))
PY
done
