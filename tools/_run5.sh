cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
P="timeout 300 python tools/probe_spmv_ab.py"
L=$PWD/legionsolvers_b200/lib
for w in c3 c4; do
 for v in la0 dbg1 dbg2 dbg1s6; do
  LSK_LIB_PATH=$L/liblsk_$v.so $P $w --ndot 1 >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$L/liblsk_$v.so $P $w --ndot 0 >> $O/ab.jsonl 2>>$O/ab.err
 done
done
LSK_LIB_PATH=$L/liblsk_la0.so $P c3 --ndot 0 >> $O/ab.jsonl 2>>$O/ab.err
LSK_LIB_PATH=$L/liblsk_la0.so $P c3 --ndot 1 >> $O/ab.jsonl 2>>$O/ab.err
python - <<'PY'
import json
for l in open("gpurun_out/r2e/ab.jsonl"):
    d=json.loads(l); print(d["workload"], d["ndot"], d["ms"], d["frac_6535"], d["env"].get("LSK_LIB_PATH","").split("_")[-1])
PY
tail -5 $O/ab.err
