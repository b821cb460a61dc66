timeout 600 python -m pytest tests/test_gpu_zz_training.py tests/test_gpu_ops.py -q -m gpu -x --tb=short -k "graphed or single_pass or bucket" 2>&1 | grep -v "^$" | tail -8 | cut -c1-700
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_c2b.json 2> gpurun_out/bench_r2_c2b.err; tail -c 400 gpurun_out/bench_r2_c2b.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_c2b.json').read().strip().splitlines()[-1]);print('c2 f16x1', d['ms_per_step'],d['value'],d['e2e']['value'],d['loss'],d['config']['step_launch'])"
timeout 300 python bench.py --config c2 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('c2 fp32', d['ms_per_step'],d['value'],d['e2e']['value'],d['loss'])"
