mkdir -p gpurun_out
for rep in 1 2; do
for w in 0 1; do
  echo "cfg2 whole=$w" >> gpurun_out/r02_whole_ab.log
  BSG_PMVT_WHOLE=$w timeout 200 python tools/profile_pmv.py --n 50000 --m 500000 --layout snp --side x --reps 20 2>&1 | tail -1 >> gpurun_out/r02_whole_ab.log
  echo "shard whole=$w" >> gpurun_out/r02_whole_ab.log
  BSG_PMVT_WHOLE=$w timeout 200 python tools/profile_pmv.py --n 487000 --m 137500 --layout snp --side x --reps 12 2>&1 | tail -1 >> gpurun_out/r02_whole_ab.log
done
done
cat gpurun_out/r02_whole_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -x -q -m gpu -k "proj or rowSums or shim or pcadapt" > gpurun_out/r02_pytest18.log 2>&1; tail -3 gpurun_out/r02_pytest18.log
timeout 300 python tools/bench_proj.py > gpurun_out/r02h_bench_proj.log 2>&1; grep "prod_and_rowSumsSq\|multLinReg" gpurun_out/r02h_bench_proj.log | cut -c1-200
