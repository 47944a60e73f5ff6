set -x
mkdir -p gpurun_out/r02
# (a) bench N=1, both arms
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_reference_arm.json 2> gpurun_out/r02/bench_reference_arm.err
# (b) launch list of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline > gpurun_out/r02/ncu_launch.log 2>&1
# (c) full captures of the two assembly kernels at full size (launches 33.. are the device-resident steps)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dg_volume_kernel --launch-skip 40 --launch-count 1 -o gpurun_out/r02/prof_volume -f python bench.py --steps 1 --warmup 3 --no-solve --no-cpu-baseline --no-graph > gpurun_out/r02/ncu_v.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dg_face_kernel --launch-skip 40 --launch-count 1 -o gpurun_out/r02/prof_face -f python bench.py --steps 1 --warmup 3 --no-solve --no-cpu-baseline --no-graph > gpurun_out/r02/ncu_f.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dg_face_sigma --launch-skip 8 --launch-count 1 --csv --log-file gpurun_out/r02/sigma.csv python bench.py --steps 1 --warmup 3 --no-solve --no-cpu-baseline --no-graph > gpurun_out/r02/ncu_s.log 2>&1
# (d) config 3b: p = 3
timeout 600 python bench.py --config 3 --degree 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_config3b.json 2> gpurun_out/r02/bench_config3b.err
timeout 600 ncu --set full --clock-control none -k regex:"dg_(volume|face)_kernel" --launch-skip 12 --launch-count 2 -o gpurun_out/r02/prof_p3 -f python bench.py --config 3 --degree 3 --steps 1 --warmup 3 --no-solve --no-cpu-baseline --no-graph > gpurun_out/r02/ncu_p3.log 2>&1
# (e) solver launch list (config 5, one solve)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/solve_launches.csv -c 6000 python tools/solve_once.py 1048576 adr 2 > gpurun_out/r02/ncu_solve.log 2>&1
# (f) other configs
for c in 1 2 3 4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_config$c.json 2> gpurun_out/r02/bench_config$c.err; done
ls -la gpurun_out/r02
