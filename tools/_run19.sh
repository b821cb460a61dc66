timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x --tb=short -k "deconv or (test_conv3d and auto)" 2>&1 | grep -v "^$" | tail -12
python tools/time_deconv.py
DA_DECONV_MMA=0 python tools/time_deconv.py
DA_ONLY="1->" timeout 300 python tools/layer_times.py | tail -2
DA_ONLY="1+1" timeout 300 python tools/layer_times.py | tail -2
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_zz_training.py tests/test_gpu_extra.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_h.json 2> gpurun_out/bench_r2_h.err; tail -c 300 gpurun_out/bench_r2_h.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_h.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
