cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n; mkdir -p $O
timeout 600 python -m pytest tests/test_host_gpu.py -m gpu -q -k "gmres_real or rmatvec" > $O/test_new.log 2>&1; echo "new tests rc=$?"; tail -3 $O/test_new.log
for cfg in "c3" "c3 --spaces 2 --steps 10" "c4 --steps 5" "c5 --steps 3" "c2 --steps 5"; do
  name=$(echo $cfg | tr ' ' '_' | tr -d '-')
  timeout 1200 python bench.py --workload $cfg > $O/bench_$name.log 2>$O/bench_$name.err; echo "bench $cfg rc=$?"; tail -2 $O/bench_$name.err
done
timeout 600 python bench.py --impl cusparse --steps 2 > $O/bench_cusparse.log 2>&1; echo "cusparse arm rc=$?"; cut -c1-600 $O/bench_cusparse.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_c*.log")):
    if "cusparse" in f: continue
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"no line",e); continue
    p=d["parity"] or {}
    print(f.split("/")[-1], d["metric"], round(d["value"],1), "roof", round(d["roofline"]["frac"],4), round(d["roofline"]["achieved"]), "GB/s iter", round(d["config"]["iteration_roofline"]["frac_of_peak"],4),
          "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["e2e"] and round(d["e2e"]["ratio_to_resident"],3), "parity", p.get("ok"), p.get("hist_rel_err"), p.get("x_rel_err"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2))
    v=d.get("vs_cusparse")
    if v: print("    vs_cusparse", {k:(round(x,3) if isinstance(x,float) else x) for k,x in v.items() if k in ("spmv_ratio","iteration_ratio","cusparse_spmv_ms","cusparse_iteration_ms","result_max_rel_diff_to_ours","unavailable")})
PY
