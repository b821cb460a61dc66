set -x
P=build/probes/umma16_probe
for fmt in 0 1; do for sh in 0 1 42; do $P check $fmt 144 $sh; done; done
$P check 0 256 3
$P rate 0 256 4000 8 1 2304 4096 1
$P rate 0 256 4000 8 2 2304 4096 4
$P rate 0 144 4000 8 2 2304 4096 4
$P rate 1 256 4000 8 2 2304 4096 4
$P rate 1 144 4000 8 2 2304 4096 4
$P rate 1 144 4000 8 3 5504 3840 6
$P rate 2 256 4000 8 2 2304 4096 4
$P rate 1 128 4000 8 2 2304 4096 4
$P rate 1 64 4000 8 2 2304 4096 4
DA_UMMA_DEBUG=1 DA_SHAPE=16,0,16,160,192,160 python tools/profile_conv.py
DA_UMMA_DEBUG=1 DA_SHAPE=32,16,16,160,192,160 python tools/profile_conv.py
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err; tail -c 600 gpurun_out/bench_r2_base.json
