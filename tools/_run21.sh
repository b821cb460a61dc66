timeout 1500 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_j.json 2> gpurun_out/bench_r2_j.err; tail -c 300 gpurun_out/bench_r2_j.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_j.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_j.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log
