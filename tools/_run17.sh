timeout 900 python -m pytest tests/test_gpu_extra.py -q -m gpu -x --tb=short -k "head_softmax or softmax_dice" 2>&1 | grep -v "^$" | tail -15
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_zz_training.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; tail -c 300 gpurun_out/bench_r2_f.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_f.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
DA_JOINT_UNFUSED=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('unfused head:',d['ms_per_step'],d['loss'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_f.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log
