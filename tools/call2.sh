set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resample" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant > gpurun_out/bench_6.json 2> gpurun_out/bench_6.err; tail -3 gpurun_out/bench_6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_6.json')); print(d['value'], d['ms_per_step']); print(d['roofline']['stage_ms_per_step'])"
for B in 64 128; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --batch-cubes $B --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B', d['config']['batch_cubes'], d['value'], d['ms_per_step']); print(d['roofline']['stage_ms_per_step'])"; done
