set -x
mkdir -p gpurun_out
for pfd in 0 2 4 8 16 32; do
  BSG_PMV_PFD=$pfd python bench.py --no-svd --no-cpu --steps 30 --warmup 5 > gpurun_out/pfd_$pfd.json 2> gpurun_out/pfd_$pfd.err
done
python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
grep -h -o '"frac": [0-9.]*' gpurun_out/pfd_*.json
cat gpurun_out/ref_arm.json
