set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_v_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs "" --gmres-m 0 > gpurun_out/r2_v_launches_bench.log 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:shell4_mma -s 2 -c 1 -o gpurun_out/r2v_shell4 python scripts/quick_perf.py quad4 > /dev/null 2>&1
$N -k regex:gather_blocks36 -s 2 -c 1 -o gpurun_out/r2v_gather36 python scripts/quick_perf.py quad4 > /dev/null 2>&1
$N -k regex:spmv6 -s 3 -c 1 -o gpurun_out/r2v_spmv6 python scripts/quick_perf.py quad4 > /dev/null 2>&1
$N -k regex:shell9_mma -s 2 -c 1 -o gpurun_out/r2v_shell9 python scripts/quick_perf.py quad9 > /dev/null 2>&1
TACSB200_OVERLAP_KINDS=0 $N -k "regex:solid_element_kernel|gather_blocks9" -s 4 -c 2 -o gpurun_out/r2v_hex8 python scripts/quick_perf.py hex8 > /dev/null 2>&1
$N -k "regex:solid_element_kernel|gather_blocks9" -s 4 -c 2 -o gpurun_out/r2v_hex27 python scripts/quick_perf.py hex27 > /dev/null 2>&1
$N -k regex:spmv3 -s 3 -c 1 -o gpurun_out/r2v_spmv3_hex27 python scripts/spmv_perf.py hex27 > /dev/null 2>&1
ls -la gpurun_out/r2v_* gpurun_out/r2_v_*
