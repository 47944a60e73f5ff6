# round 2, session 3: solver kernels A/B (block-row SpMV, unrolled multigrid block rows, parallel dense inverse) + evidence
set -x
mkdir -p gpurun_out/r02b
lscpu | grep -i "model name\|^CPU(s)\|numa\|thread\|socket" > gpurun_out/r02b/host.txt; nproc >> gpurun_out/r02b/host.txt
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_convergence.py -x -q > gpurun_out/r02b/solver_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/solver_tests.log
tail -3 gpurun_out/r02b/solver_tests.log
for v in "" build/lib_spmv_old.so build/lib_solv_old.so "" build/lib_solv_old.so; do
  if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
  echo "lib=$v" >> gpurun_out/r02b/solve_ab.log
  timeout 300 python tools/solve_once.py 1048576 adr 2 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab.log
  timeout 300 python tools/solve_once.py 1048576 diffusion 2 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab.log
done
unset REYNA_B200_LIB
cat gpurun_out/r02b/solve_ab.log
for c in 1 2; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b/bench_config$c.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/solve_launches.csv -c 6000 python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_blockrow_kernel|mgb_block_row" --launch-skip 30 --launch-count 3 -o gpurun_out/r02b/prof_solver -f python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solver_full.log 2>&1
ls -la gpurun_out/r02b
