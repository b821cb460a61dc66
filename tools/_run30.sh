timeout 1500 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | grep -E "^E |passed|failed|Error" | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-cuda --no-parity > gpurun_out/bench_r2_o.json 2> gpurun_out/bench_r2_o.err; tail -c 300 gpurun_out/bench_r2_o.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_o.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['e2e']['value'],d['loss'],d['gpu_launches'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_o.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -1 gpurun_out/prof_step.log
