set -x
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_5.json 2> gpurun_out/bench_5.err; tail -3 gpurun_out/bench_5.err
python tools/exp_l2gran.py 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 --no-variant --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'postproc_stitch|gather3|prefilter|select_hist|extract_tma|normalize_apply' --launch-skip 0 -c 14 -o gpurun_out/prof_r01b python bench.py --steps 1 --warmup 1 --no-variant --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
