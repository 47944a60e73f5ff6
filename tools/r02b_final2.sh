# final verification of round 2 (after the last code changes): full GPU suite, smoke, main bench line, p = 3 line
set -x
O=gpurun_out/r02g
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
tail -4 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --config 3 --degree 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config3b.json 2> $O/bench_config3b.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
