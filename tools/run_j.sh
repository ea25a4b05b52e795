mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scaling_reuse or degenerate or prodvec or prodVec" > gpurun_out/r02_pytest20.log 2>&1; tail -3 gpurun_out/r02_pytest20.log
timeout 300 python bench.py --workload cfg2 --no-cpu --no-extra --no-svd --steps 20 --warmup 5 > gpurun_out/r02j_bench_cfg2.json 2> gpurun_out/r02j.err; tail -3 gpurun_out/r02j.err
python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench_cfg2.json').read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['roofline']['frac']); e=d['e2e']; print({k:e[k] for k in e if k!='note'})"
