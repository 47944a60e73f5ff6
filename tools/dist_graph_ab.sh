# A/B of the distributed solve with and without CUDA-graph capture of the preconditioner application (NCCL calls inside)
N=${1:-2}
for g in 0 1; do
  echo "== REYNA_B200_DIST_GRAPHS=$g"
  REYNA_B200_DIST_GRAPHS=$g timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+g)) bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['solve']; print('wall', round(s['dgfem_solve_true_wall_s'],3), 'solve_s', round(s['solve_s'],3), 'setup', round(s['setup_s'],3), 'iterations', s['iterations'], 'l2', s['l2_error_vs_exact'])"
done
