set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_pytest11.log 2>&1; tail -5 gpurun_out/r02_pytest11.log
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "ksplit" > gpurun_out/r02_pytest12.log 2>&1; tail -3 gpurun_out/r02_pytest12.log
for r in 0.002 0.01 0.03; do
  timeout 300 python bench.py --workload cfg2 --layout snp --na-rate $r --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02c_bench_cfg2_snp_na$r.json 2> gpurun_out/r02c_na$r.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02c_bench_cfg2_snp_na$r.json').read().strip().splitlines()[-1])
print('na $r', d['ms_per_step'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'), d.get('svd',{}).get('nops'))"
done
timeout 300 python bench.py --workload cfg2 --layout both --na-rate 0.01 --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02c_bench_cfg2_both_na0.01.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_cfg2_both_na0.01.json').read().strip().splitlines()[-1]); print('both 1%', d['ms_per_step'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'))"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_corr -s 2 -c 1 -o gpurun_out/r02_kcorr_after2 -f python tools/profile_pmv.py --n 50000 --m 500000 --na-rate 0.01 --layout snp --side x --reps 4 > gpurun_out/r02_kcorr_after2_ncu.log 2>&1; tail -2 gpurun_out/r02_kcorr_after2_ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_na1pct.csv python tools/profile_pmv.py --n 50000 --m 500000 --na-rate 0.01 --layout snp --side both --reps 3 > /dev/null 2>&1
