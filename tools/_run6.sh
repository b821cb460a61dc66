rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_zz_training.py -q -m gpu 2>&1 | tail -25
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 300 gpurun_out/bench_r2_a.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_a.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
DA_CONV_SPLIT=tf32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_a_tf32.json 2>/dev/null;  python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_a_tf32.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r2_a.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -3 gpurun_out/prof_step.log
