set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02_pytest16.log 2>&1; tail -3 gpurun_out/r02_pytest16.log
for r in 0.0 0.01; do
  timeout 300 python bench.py --workload cfg2 --layout snp --na-rate $r --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02f_bench_cfg2_snp_na$r.json 2> gpurun_out/r02f_na$r.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02f_bench_cfg2_snp_na$r.json').read().strip().splitlines()[-1])
print('na $r', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'), d.get('svd',{}).get('nops'), d['e2e']['value'])"
done
