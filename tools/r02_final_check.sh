set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_final_tests.log
tail -5 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 3000 gpurun_out/r2_bench_n1.json
REYNA_B200_MG_FP32=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_fp64mg.json 2> gpurun_out/r2_bench_n1_fp64mg.err; tail -c 1500 gpurun_out/r2_bench_n1_fp64mg.json
