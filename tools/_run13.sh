for i in 1 2 3 4 5 6; do timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k test_conv3d --tb=line 2>&1 | grep -v "^$" | tail -2; done
for sh in 16,0,16,160,192,160 32,0,16,160,192,160 32,16,16,160,192,160 16,0,48,160,192,160 32,0,32,80,96,80 64,32,32,80,96,80 64,64,64,40,48,40; do
  DA_SHAPE=$sh timeout 120 python tools/time_conv.py
  DA_UMMA_TMA_IN=0 DA_SHAPE=$sh timeout 120 python tools/time_conv.py
done
DA_SHAPE=32,0,16,160,192,160 DA_NT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel -s 2 -c 1 -o gpurun_out/umma_tma_in -f python tools/time_conv.py > gpurun_out/ncu13.log 2>&1; tail -3 gpurun_out/ncu13.log
