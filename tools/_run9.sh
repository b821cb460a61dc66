timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k test_conv3d --tb=line 2>&1 | grep -v "^$" | tail -12
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_nets.py -q -m gpu --tb=line 2>&1 | grep -v "^$" | tail -20
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -c 300 gpurun_out/bench_r2_b.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_b.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
