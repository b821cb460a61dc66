timeout 900 python -m pytest tests/test_gpu_zz_training.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -12
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_l.json 2> gpurun_out/bench_r2_l.err; tail -c 600 gpurun_out/bench_r2_l.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_l.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'],d['gpu_launches'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda --no-graph 2> /dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('eager:',d['ms_per_step'],d['value'],d['e2e'],d['loss'],d['gpu_launches'])"
