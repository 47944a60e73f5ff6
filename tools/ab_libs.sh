# A/B of library builds on one box: REYNA_B200_LIB selects the .so (variants built with extra -D flags into build/)
for v in "" build/lib_oldv.so "" build/lib_oldv.so; do
  if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-solve --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
done
