show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],3), round(d["ms_per_step"],3)); print(d["roofline"]["stage_ms_per_step"])'
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 2>/dev/null | python -c "$show"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stitch" 2>&1 | tail -2
