set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputest_final.txt 2>&1; tail -3 gpurun_out/r02_gputest_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 600 python tools/bench_extra.py --cor-m 200000 --clump-n 100000 --clump-m 200000 --grm-m 1000000 > gpurun_out/r02_extra_final.jsonl 2> gpurun_out/r02_extra_final.err; cat gpurun_out/r02_extra_final.jsonl | cut -c1-400
timeout 900 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; tail -c 600 gpurun_out/r02_bench_final_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err; tail -c 400 gpurun_out/r02_bench_final_ref.json
