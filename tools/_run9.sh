cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
B="timeout 300 python bench.py --no-parity --no-e2e --no-cusparse --no-cpu-baseline --steps 20"
run() { # name, env..., -- args
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs $B "$@" > $O/$name.log 2>$O/$name.err || echo "$name failed"
  python - "$O/$name.log" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"],1), "it/s", round(1e6/d["value"],2), "us/it", "spmv", round(d["roofline"]["ms_per_launch"]*1e3,2), "us", "launches", d["gpu_launches"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
for shape in "32,256,256" "256,256,256"; do
 for tr in graph eager; do
  for pdl in 0 1 4 5 7; do
    run "s${shape%%,*}_${tr}_pdl${pdl}" LSK_PDL=$pdl LSK_TRACE=$tr -- --shape $shape
  done
 done
done
# slab with the TMA-streamed update kernel (PDL-capable) instead of the grid-stride one
for pdl in 0 7; do
  run "s32_graph_tmaupd_pdl${pdl}" LSK_PDL=$pdl LSK_TMA_STREAM_MIN_PACKS=1000 -- --shape 32,256,256
done
