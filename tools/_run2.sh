mkdir -p gpurun_out/r2b
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "spmv" -x -q > $O/test_spmv.log 2>&1; echo "spmv tests rc=$?"
tail -3 $O/test_spmv.log
P="timeout 300 python tools/probe_spmv_ab.py"
for w in c3 c2 c4 slab8; do
  LSK_SPMV_IMPL=tma $P $w --ndot 1 --save /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  LSK_SPMV_DYN=1 $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$PWD/legionsolvers_b200/lib/liblsk_s6b1.so $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$PWD/legionsolvers_b200/lib/liblsk_s2b3.so LSK_WS_CTAS=3 $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
  LSK_LIB_PATH=$PWD/legionsolvers_b200/lib/liblsk_s2b3.so LSK_WS_CTAS=3 LSK_SPMV_DYN=1 $P $w --ndot 1 --check /tmp/ref_$w.pt >> $O/ab.jsonl 2>>$O/ab.err
done
$P c3 --ndot 0 >> $O/ab.jsonl 2>>$O/ab.err
python - <<'PY'
import json
for l in open("gpurun_out/r2b/ab.jsonl"):
    d=json.loads(l); print(d["workload"], d["ndot"], d["ms"], d["frac_6535"], d.get("y_bit_identical"), d["env"])
PY
tail -5 $O/ab.err
