cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "spmv" -x -q > $O/test_spmv.log 2>&1; echo "spmv tests rc=$?"
tail -5 $O/test_spmv.log
P="timeout 300 python tools/probe_spmv_ab.py"
L=$PWD/legionsolvers_b200/lib
for w in c3 c2 c4 slab8; do
  LSK_SPMV_IMPL=tma $P $w --ndot 1 --save /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  $P $w --ndot 0 >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$L/liblsk_s2b3.so LSK_WS_CTAS=3 $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$L/liblsk_dbg1.so $P $w --ndot 1 >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$L/liblsk_dbg1s6.so $P $w --ndot 1 >> $O/ab.jsonl 2>>$O/ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2f/ab.jsonl"):
    d=json.loads(l); print(d["workload"], d["ndot"], d["ms"], d["frac_6535"], d.get("y_bit_identical"), d["env"].get("LSK_LIB_PATH","").split("_")[-1], d["env"].get("LSK_SPMV_IMPL",""))
PY
tail -5 $O/ab.err
