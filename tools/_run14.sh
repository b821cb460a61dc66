for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k test_conv3d --tb=line 2>&1 | grep -v "^$" | tail -2; done
for sh in 16,0,16,160,192,160 32,0,16,160,192,160 32,16,16,160,192,160 16,0,48,160,192,160 32,0,32,80,96,80 64,32,32,80,96,80 64,64,64,40,48,40; do
  DA_SHAPE=$sh timeout 120 python tools/time_conv.py
done
DA_UMMA_DEBUG=1 DA_SHAPE=32,0,16,160,192,160 timeout 120 python tools/time_conv.py
DA_UMMA_DEBUG=1 DA_SHAPE=16,0,16,160,192,160 timeout 120 python tools/time_conv.py
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_zz_training.py -q -m gpu --tb=line 2>&1 | grep -v "^$" | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; tail -c 300 gpurun_out/bench_r2_d.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_d.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r2_d.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log
