timeout 900 python -m pytest tests/test_gpu_zz_training.py tests/test_gpu_nets.py tests/test_gpu_golden.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -8
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_extra.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_k.json 2> gpurun_out/bench_r2_k.err; tail -c 300 gpurun_out/bench_r2_k.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_k.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'],d['gpu_launches'])"
DA_IMPL=3 DA_ONLY="enc3" timeout 300 python tools/layer_times.py | tail -3
DA_IMPL=3 DA_ONLY="reg.dec1" timeout 300 python tools/layer_times.py | tail -2
DA_IMPL=3 DA_ONLY="reg.dec2" timeout 300 python tools/layer_times.py | tail -2
