#!/bin/bash
# Round profile: (1) plain bench (must exit 0 first), (2) ncu launch list of the same command,
# (3) one `ncu --set full` capture of the hot kernels.  Numbers printed under ncu are not bench values.
set -x
TAG=${1:-r02}
SHORT="--no-variant --no-cpu-baseline --no-strong --no-config5 --no-dropin --e2e-steps 1"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 $SHORT \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'postproc_stitch|march_yz|cols_reg|rows_pipe|rows_reg|select_guided|select_hist|normalize_apply|extract_tma|af3_fill|af3_bin' -c 26 \
    -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 0 $SHORT \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
