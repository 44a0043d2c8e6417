timeout 600 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["af3_mode"], "B", d["config"]["batch_cubes"], round(d["value"],3), round(d["ms_per_step"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3)); print(d["roofline"]["stage_ms_per_step"])'
for B in 32 128; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --batch-cubes $B --e2e-steps 1 2>/dev/null | python -c "$show"; done
for P in 0 1 2 3; do echo promo $P; MICA_TMA_L2PROMO=$P python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variant --af3-mode dense --batch-cubes 32 --e2e-steps 1 2>/dev/null | python -c "$show"; done
MICA_TMA_L2PROMO=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variant --af3-mode dense --batch-cubes 128 --e2e-steps 1 2>/dev/null | python -c "$show"
