timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k test_conv3d --tb=line 2>&1 | grep -v "^$" | tail -4
for sh in 16,0,16,160,192,160 32,0,16,160,192,160 32,16,16,160,192,160 32,0,32,80,96,80 64,32,32,80,96,80 64,64,64,40,48,40 16,0,48,160,192,160; do
  DA_SHAPE=$sh timeout 120 python tools/profile_conv.py
  DA_UMMA_TMA_IN=0 DA_SHAPE=$sh timeout 120 python tools/profile_conv.py
  DA_UMMA_DEBUG=1 DA_SHAPE=$sh timeout 120 python tools/profile_conv.py | tail -1
done
