cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
python tools/probe_spmv_ab.py c3 --ndot 1 --reps 5 > $O/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:csr_ws_kernel -s 6 -c 1 -o $O/ws_c3 python tools/probe_spmv_ab.py c3 --ndot 1 --reps 5 > $O/ncu_c3.log 2>&1
echo "c3 rc=$?"
python tools/probe_spmv_ab.py c4 --ndot 1 --reps 5 > $O/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:csr_ws_kernel -s 6 -c 1 -o $O/ws_c4 python tools/probe_spmv_ab.py c4 --ndot 1 --reps 5 > $O/ncu_c4.log 2>&1
echo "c4 rc=$?"
tail -3 $O/ncu_c3.log $O/ncu_c4.log
ls -la $O
