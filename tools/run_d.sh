set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "missing or na or NA or prodvec or svd or SVD" > gpurun_out/r02_pytest13.log 2>&1; tail -3 gpurun_out/r02_pytest13.log
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "ksplit" > gpurun_out/r02_pytest14.log 2>&1; tail -3 gpurun_out/r02_pytest14.log
for r in 0.002 0.01 0.03; do
  timeout 300 python bench.py --workload cfg2 --layout snp --na-rate $r --no-cpu --no-extra --steps 20 --warmup 5 > gpurun_out/r02d_bench_cfg2_snp_na$r.json 2> gpurun_out/r02d_na$r.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02d_bench_cfg2_snp_na$r.json').read().strip().splitlines()[-1])
print('na $r', d['ms_per_step'], d['roofline']['frac'], d.get('svd',{}).get('wall_s'), d.get('svd',{}).get('nops'))"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_corr -s 2 -c 1 -o gpurun_out/r02_kcorr_after3 -f python tools/profile_pmv.py --n 50000 --m 500000 --na-rate 0.01 --layout snp --side x --reps 4 > gpurun_out/r02_kcorr_after3_ncu.log 2>&1; tail -2 gpurun_out/r02_kcorr_after3_ncu.log
