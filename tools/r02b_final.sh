# round 2 final evidence on one B200 (everything lands in gpurun_out/r02f; the summaries are copied to profiles/ afterwards)
set -x
O=gpurun_out/r02f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
tail -4 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 1200 $O/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/solve_launches.csv -c 6000 python tools/solve_once.py 1048576 adr 2 > $O/ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_blockrow_kernel|mgb_block_row" --launch-skip 30 --launch-count 3 -o $O/prof_solver -f python tools/solve_once.py 1048576 adr 2 > $O/ncu_solver_full.log 2>&1
for c in 1 2 3 4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config$c.json 2> $O/bench_config$c.err; done
timeout 600 python bench.py --config 3 --degree 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config3b.json 2> $O/bench_config3b.err
ls -la $O

# ---- second stage (after the last code changes of the round): bash tools/r02b_final.sh is run again with O=gpurun_out/r02g;
#      the short form that was used:
#   O=gpurun_out/r02g
#   mkdir -p $O
#   timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
#   tail -4 $O/gpu_tests.log
#   timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
#   timeout 600 python bench.py --config 3 --degree 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_config3b.json 2> $O/bench_config3b.err
#   timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
