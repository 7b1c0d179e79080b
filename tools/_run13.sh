cd $GRAFT_REPO_ROOT
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi -L > $O/test_gpu_2gpus.log
timeout 2400 python -m pytest tests -m gpu -q -rs --durations=5 >> $O/test_gpu_2gpus.log 2>&1; echo "gpu tests rc=$?"
tail -22 $O/test_gpu_2gpus.log
