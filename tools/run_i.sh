mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -x -q -m gpu > gpurun_out/r02_pytest19.log 2>&1; tail -3 gpurun_out/r02_pytest19.log
timeout 300 python tools/bench_proj.py > gpurun_out/r02i_bench_proj.log 2>&1; grep "prod_and_rowSumsSq\|multLinReg\|cpu_oracle" gpurun_out/r02i_bench_proj.log | cut -c1-420
