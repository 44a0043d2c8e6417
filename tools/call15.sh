show='import json,sys; d=json.loads(sys.stdin.read()); print(d["af3_mode"], round(d["value"],3), round(d["ms_per_step"],3)); print(d["roofline"]["stage_ms_per_step"])'
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 2>/dev/null | python -c "$show"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 --af3-mode dense 2>/dev/null | python -c "$show"
