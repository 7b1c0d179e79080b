cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi -L | wc -l
for extra in "X=1" "LSK_HALO_OPEN=0" "LSK_PDL=0"; do
 env $extra timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --no-cpu-baseline > $O/bench_n8_$extra.log 2>$O/bench_n8_$extra.err; echo "bench n8 [$extra] rc=$?"
 tail -2 $O/bench_n8_$extra.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j/bench_n8_*.log")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"no line",e); continue
    print(f, round(d["value"],1), "it/s", round(1e6/d["value"],2), "us/it", "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ratio_to_resident"],3), "parity", d["parity"]["ok"], d["parity"]["hist_rel_err"], d["parity"]["x_rel_err"])
    print("   ", d["config"]["time_inside_collectives"], d["config"]["spmv_ms_per_launch_by_rank"], d["config"]["comm_error"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
