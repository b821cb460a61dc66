timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x --tb=short -k "test_conv3d and auto" 2>&1 | grep -v "^$" | tail -5
DA_ONLY="1->" timeout 300 python tools/layer_times.py | tail -2
DA_ONLY="1+1" timeout 300 python tools/layer_times.py | tail -2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_i.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log
