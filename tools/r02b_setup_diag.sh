set -x
O=gpurun_out/r02f; mkdir -p $O
for i in 1 2 3; do timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('lazy $i', s['dgfem_solve_true_wall_s'], s['solve_s'], s['setup_s'])" >> $O/setup_diag.log; done
for i in 1 2; do CUDA_MODULE_LOADING=EAGER timeout 300 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('eager $i', s['dgfem_solve_true_wall_s'], s['solve_s'], s['setup_s'])" >> $O/setup_diag.log; done
cat $O/setup_diag.log
