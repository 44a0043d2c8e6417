ncu --set full --clock-control none --import-source on -k regex:'march3|prefilter|select_hist|postproc_stitch|normalize_apply' -c 9 -o gpurun_out/prof_r01c python bench.py --steps 1 --warmup 0 --no-variant --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_e.log 2>&1
tail -3 gpurun_out/ncu_e.log
