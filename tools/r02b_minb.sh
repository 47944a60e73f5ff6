set -x
O=gpurun_out/r02f
mkdir -p $O
for v in "" build/lib_mgb_minb6.so build/lib_mgb_minb8.so ""; do
  if [ -n "$v" ]; then export REYNA_B200_LIB=$PWD/$v; else unset REYNA_B200_LIB; fi
  timeout 300 python tools/solve_repeat.py 1048576 adr 2 5 2>/dev/null | tail -1 >> $O/solve_minb.log
done
cat $O/solve_minb.log
