mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_shim.py -x -q -m gpu -k "clump or autoSVD or shim or cfg3 or dosage" > gpurun_out/r02_pytest21.log 2>&1; tail -3 gpurun_out/r02_pytest21.log
timeout 400 python tools/bench_extra.py --cor-m 20000 --clump-n 100000 --clump-m 200000 --skip grm > gpurun_out/r02k_extra.jsonl 2> gpurun_out/r02k_extra.err; grep clumping gpurun_out/r02k_extra.jsonl | cut -c1-400
BSG_CLUMP_HOST=1 timeout 400 python tools/bench_extra.py --cor-m 20000 --clump-n 100000 --clump-m 200000 --skip grm 2>/dev/null | grep clumping | cut -c1-300
