set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "missing or na or NA or prodvec or svd or SVD" > gpurun_out/r02_pytest15.log 2>&1; tail -3 gpurun_out/r02_pytest15.log
for r in 0.002 0.01 0.03; do
  timeout 300 python bench.py --workload cfg2 --layout snp --na-rate $r --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02e_bench_cfg2_snp_na$r.json 2> gpurun_out/r02e_na$r.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02e_bench_cfg2_snp_na$r.json').read().strip().splitlines()[-1])
print('na $r', d['ms_per_step'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'), d.get('svd',{}).get('nops'))"
done
timeout 300 python bench.py --workload cfg2 --layout snp --na-rate 0.0 --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02e_bench_cfg2_snp_na0.json 2>/dev/null
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02e_bench_cfg2_snp_na0.json').read().strip().splitlines()[-1])
print('na 0', d['ms_per_step'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'), d.get('svd',{}).get('nops'))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02e_launches_na1pct.csv python tools/profile_pmv.py --n 50000 --m 500000 --na-rate 0.01 --layout snp --side both --reps 3 > /dev/null 2>&1
grep -c k_corr gpurun_out/r02e_launches_na1pct.csv
