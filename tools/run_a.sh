set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rds_golden or clumping" > gpurun_out/r02_pytest8.log 2>&1; tail -3 gpurun_out/r02_pytest8.log
for shape in "487000 550000" "487000 275000"; do
  set -- $shape
  for ks in 0 5 6 8 10 12 16 20; do
    echo "shape $1 x $2 ks $ks" >> gpurun_out/r02_ks_sweep2.log
    BSG_PMVT_KS=$ks timeout 200 python tools/profile_pmv.py --n $1 --m $2 --layout snp --side x --reps 10 2>&1 | tail -1 >> gpurun_out/r02_ks_sweep2.log
  done
done
cat gpurun_out/r02_ks_sweep2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_corr -s 2 -c 1 -o gpurun_out/r02_kcorr -f python tools/profile_pmv.py --n 50000 --m 500000 --na-rate 0.01 --layout snp --side x --reps 4 > gpurun_out/r02_kcorr_ncu.log 2>&1; tail -2 gpurun_out/r02_kcorr_ncu.log
