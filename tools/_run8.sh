cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $O/test_multi_gpu.log 2>&1; echo "multi-gpu tests rc=$?"
tail -30 $O/test_multi_gpu.log
for extra in "" "LSK_HALO_OPEN=0"; do
 env $extra timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --no-cpu-baseline > $O/bench_n2_$extra.log 2>$O/bench_n2_$extra.err; echo "bench n2 [$extra] rc=$?"
 tail -3 $O/bench_n2_$extra.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h/bench_n2_*.log")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"no line",e); continue
    print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ratio_to_resident"],3), "parity", d["parity"]["ok"], d["parity"]["hist_rel_err"], d["config"]["time_inside_collectives"], d["config"]["spmv_ms_per_launch_by_rank"], d["config"]["comm_error"], d["config"]["collectives"])
PY
