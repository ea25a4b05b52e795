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
