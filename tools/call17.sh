timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_fullsize.py -m gpu -q -x -k "normalize or order_stat or slab or golden" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],3), round(d["ms_per_step"],3)); print(d["roofline"]["stage_ms_per_step"])'
for i in 1 2 3; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 2>/dev/null | python -c "$show"; done
