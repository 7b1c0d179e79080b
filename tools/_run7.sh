cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -8 $O/test_gpu.log
timeout 900 python bench.py > $O/bench_c3.log 2>$O/bench_c3.err; echo "bench c3 rc=$?"; tail -3 $O/bench_c3.err
timeout 900 python bench.py --workload c2 --steps 5 > $O/bench_c2.log 2>$O/bench_c2.err; echo "bench c2 rc=$?"; tail -3 $O/bench_c2.err
timeout 900 python bench.py --workload c4 --steps 5 > $O/bench_c4.log 2>$O/bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/bench_c4.err
python - <<'PY'
import json
for w in ("c3","c2","c4"):
    try:
        d=json.loads(open(f"gpurun_out/r2g/bench_{w}.log").read().strip().splitlines()[-1])
    except Exception as e:
        print(w, "no line", e); continue
    print(w, d["metric"], round(d["value"],1), "roof", round(d["roofline"]["frac"],4), "iter", round(d["config"]["iteration_roofline"]["frac_of_peak"],4),
          "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["e2e"] and round(d["e2e"]["ratio_to_resident"],3), "parity", d["parity"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "cusp", d["vs_cusparse"])
PY
