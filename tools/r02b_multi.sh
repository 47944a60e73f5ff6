# multi-GPU check of the round-2 final code: bash tools/r02b_multi.sh N   (N = 2: also the two-process NCCL tests)
set -x
N=${1:-2}
O=gpurun_out/r02f
mkdir -p $O
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q > $O/two_gpu_tests.log 2>&1; echo "rc=$?" >> $O/two_gpu_tests.log; tail -3 $O/two_gpu_tests.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 1500 $O/bench_n$N.json
