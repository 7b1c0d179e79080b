set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L
cd $GRAFT_REPO_ROOT
# (a) the new kernel against the oracle: SpMV tests only
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "spmv" -x -q > gpurun_out/r2a/test_spmv.log 2>&1; echo "spmv tests rc=$?" 
tail -5 gpurun_out/r2a/test_spmv.log
# (b) A/B probes
P="timeout 300 python tools/probe_spmv_ab.py"
for w in c3 c2 c4 slab8; do
  LSK_SPMV_IMPL=tma $P $w --ndot 1 --save /tmp/ref_$w.pt >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
  $P $w --ndot 1 --check /tmp/ref_$w.pt >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
  LSK_SPMV_DYN=1 $P $w --ndot 1 --check /tmp/ref_$w.pt >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
  LSK_LIB_PATH=$PWD/legionsolvers_b200/lib/liblsk_s6b1.so $P $w --ndot 1 --check /tmp/ref_$w.pt >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
  LSK_LIB_PATH=$PWD/legionsolvers_b200/lib/liblsk_s2b3.so LSK_WS_CTAS=3 $P $w --ndot 1 --check /tmp/ref_$w.pt >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
done
LSK_SPMV_IMPL=tma $P c3 --ndot 0 >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
$P c3 --ndot 0 >> gpurun_out/r2a/ab.jsonl 2>>gpurun_out/r2a/ab.err
cat gpurun_out/r2a/ab.jsonl
tail -20 gpurun_out/r2a/ab.err
# (c) the whole GPU suite
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a/test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -15 gpurun_out/r2a/test_gpu.log
# (d) bench
timeout 600 python bench.py > gpurun_out/r2a/bench_n1.log 2>gpurun_out/r2a/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/r2a/bench_n1.log | cut -c1-1500
tail -5 gpurun_out/r2a/bench_n1.err
LSK_SPMV_IMPL=tma timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a/bench_n1_tma.log 2>&1; echo "bench tma rc=$?"
cut -c1-400 gpurun_out/r2a/bench_n1_tma.log
