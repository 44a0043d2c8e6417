timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -q -x -k "resample or slab" 2>&1 | tail -5
MICA_NO_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resample" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.read()); print(d["af3_mode"], "B", d["config"]["batch_cubes"], round(d["value"],3), round(d["ms_per_step"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3)); print(d["roofline"]["stage_ms_per_step"])'
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 1 2>/dev/null | python -c "$show"
