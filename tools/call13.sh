show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],3), round(d["ms_per_step"],3)); print(d["roofline"]["stage_ms_per_step"])'
for L in 16 32; do echo lines $L; MICA_PREFILTER_LINES=$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 2>/dev/null | python -c "$show"; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "resample or long_lines or linear" 2>&1 | tail -3
