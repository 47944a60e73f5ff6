set -x
mkdir -p gpurun_out/r02b
for v in "" build/lib_mgb_late.so build/lib_solv_old.so "" build/lib_mgb_late.so; do
  if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
  timeout 300 python tools/solve_repeat.py 1048576 adr 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
done
unset REYNA_B200_LIB
timeout 300 python tools/solve_repeat.py 1048576 adr 2 4 '{"mg_fp32": true}' 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
timeout 300 python tools/solve_repeat.py 1048576 diffusion 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
timeout 300 python tools/solve_repeat.py 65536 adr 2 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
timeout 300 python tools/solve_repeat.py 1024 adr 1 4 2>/dev/null | tail -1 >> gpurun_out/r02b/solve_ab2.log
cat gpurun_out/r02b/solve_ab2.log
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/r02b/solver_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r02b/solver_tests2.log; tail -2 gpurun_out/r02b/solver_tests2.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/solve_launches2.csv -c 3000 python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02b/ncu_solve2.log 2>&1
