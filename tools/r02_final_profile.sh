cd $GRAFT_REPO_ROOT
O=gpurun_out/r2fin; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gpu_tests.log
LSK_SPMV_IMPL=tma timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_config_size_gpu.py -m gpu -q -k "spmv or csr" > $O/gpu_tests_tma_impl.log 2>&1; echo "pytest(tma impl) rc=$?"; tail -2 $O/gpu_tests_tma_impl.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.log 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --iters-per-step 5 --no-cpu-baseline --no-parity --no-e2e --no-cusparse"
$CMD > $O/plain_bench.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file $O/r02_launch_list_raw.csv $CMD > $O/ncu_launch.log 2>&1
echo "launch list rc=$?"
for w in c3 c2 c4; do
  python tools/probe_spmv_ab.py $w --ndot 1 --reps 5 > $O/plain_$w.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:csr_ws_kernel -s 6 -c 1 -o $O/r02_ws_$w python tools/probe_spmv_ab.py $w --ndot 1 --reps 5 > $O/ncu_$w.log 2>&1
  echo "$w rc=$?"
  ncu -i $O/r02_ws_$w.ncu-rep --page raw --csv > $O/r02_ws_${w}_raw.csv 2>/dev/null
done
ncu -i $O/r02_ws_c3.ncu-rep --page source --csv > $O/r02_ws_c3_source.csv 2>/dev/null
rm -f $O/r02_ws_c2.ncu-rep $O/r02_ws_c4.ncu-rep
python tools/probe_coo.py > $O/plain_coo.log 2>&1 && \
timeout 300 ncu --set full --clock-control none -k regex:coo_segreduce -s 3 -c 1 -o $O/r02_coo python tools/probe_coo.py > $O/ncu_coo.log 2>&1
echo "coo rc=$?"
ncu -i $O/r02_coo.ncu-rep --page raw --csv > $O/r02_coo_raw.csv 2>/dev/null; rm -f $O/r02_coo.ncu-rep
$CMD > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:"cg_update_tma|cg_direction_tma" -s 20 -c 2 -o $O/r02_cg_vec $CMD > $O/ncu_cgvec.log 2>&1
echo "cg vec rc=$?"
ncu -i $O/r02_cg_vec.ncu-rep --page raw --csv > $O/r02_cg_vec_raw.csv 2>/dev/null; rm -f $O/r02_cg_vec.ncu-rep
ls -la $O
