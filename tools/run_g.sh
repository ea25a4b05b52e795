set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02_pytest17.log 2>&1; tail -3 gpurun_out/r02_pytest17.log
timeout 300 python bench.py --workload cfg2 --layout snp --no-cpu --no-extra --no-svd --steps 20 --warmup 5 > gpurun_out/r02g_bench_cfg2_snp.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02g_bench_cfg2_snp.json').read().strip().splitlines()[-1]); print('cfg2 snp', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
timeout 600 python bench.py --no-cpu --no-extra --no-svd --steps 20 --warmup 5 > gpurun_out/r02g_bench_cfg5.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02g_bench_cfg5.json').read().strip().splitlines()[-1]); print('cfg5', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
