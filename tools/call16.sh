timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_api" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variant --e2e-steps 3 2>gpurun_out/b16.err | python -c '
import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],3), round(d["ms_per_step"],3), d["e2e"])'
tail -3 gpurun_out/b16.err
