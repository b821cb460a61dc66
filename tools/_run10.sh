python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; tail -c 400 gpurun_out/bench_r2_c.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_c.json').read().strip().splitlines()[-1])
print(d['ms_per_step'],d['value'],d['e2e'],d['loss'])
for k in ('cpu_baseline','torch_cuda_baseline','parity'): print(k, d[k])
print({k:d['roofline'][k] for k in ('achieved','frac','launch_ms','tensor_pipe_occupancy','hbm_frac')}, {k:d['roofline_wgrad'][k] for k in ('achieved','frac','launch_ms')})
"
nproc; free -g | head -2
for c in c3 c5 c2; do python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-torch-cuda > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; tail -c 300 gpurun_out/bench_r2_$c.err; python -c "
import json,sys;d=json.loads(open('gpurun_out/bench_r2_$c.json').read().strip().splitlines()[-1]);print('$c',d['ms_per_step'],d['value'],d['peak_mem_gb'],d['loss'])"; done
