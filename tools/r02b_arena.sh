set -x
O=gpurun_out/r02f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_probes.py -x -q > $O/solver_tests_arena.log 2>&1; echo "rc=$?" >> $O/solver_tests_arena.log; tail -2 $O/solver_tests_arena.log
timeout 300 python tools/solve_repeat.py 1048576 adr 2 8 2>/dev/null | tail -1 >> $O/solve_arena.log
timeout 300 python tools/solve_repeat.py 16384 diffusion 1 8 2>/dev/null | tail -1 >> $O/solve_arena.log
timeout 300 python tools/solve_repeat.py 1024 adr 1 8 2>/dev/null | tail -1 >> $O/solve_arena.log
cat $O/solve_arena.log
for c in 1 2; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config$c.json 2> $O/bench_config$c.err; done
