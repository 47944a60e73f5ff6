# round 2, session 3: full GPU suite, then A/B of the volume-kernel variants (tools/build_variant.sh) and of the malloc tuning
set -x
mkdir -p gpurun_out/r02b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b/tests.log
tail -4 gpurun_out/r02b/tests.log
run() {  # name lib
  if [ -n "$2" ]; then export REYNA_B200_LIB=$PWD/$2; else unset REYNA_B200_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline > gpurun_out/r02b/bench_$1.json 2> gpurun_out/r02b/bench_$1.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02b/bench_$1.json')); print('$1', 'step', round(d['ms_per_step'],4), 'call', round(d['roofline']['kernel_ms'],4), 'e2e_s', round(d['e2e']['s_per_step'],4), d['roofline']['fp64'])"
}
run base ""
for v in pf pf184 r184 lb11; do run $v build/lib_$v.so; done
run base2 ""
unset REYNA_B200_LIB
REYNA_B200_MALLOPT=0 run nomallopt ""
run mallopt ""
