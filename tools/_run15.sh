cd $GRAFT_REPO_ROOT
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "coo or gmres or planner_matvec or rmatvec" > $O/test_coo.log 2>&1; echo "coo tests rc=$?"; tail -5 $O/test_coo.log
LSK_COO_IMPL=seg timeout 300 python tools/probe_coo.py 2>&1 | tail -3
timeout 300 python tools/probe_coo.py 2>&1 | tail -3
