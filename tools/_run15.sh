for sh in 32,0,16,160,192,160 32,16,16,160,192,160 64,32,32,80,96,80 64,64,64,40,48,40; do
  DA_SHAPE=$sh timeout 120 python tools/time_conv.py
done
DA_UMMA_DEBUG=1 DA_SHAPE=32,0,16,160,192,160 timeout 120 python tools/time_conv.py
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k test_conv3d --tb=line 2>&1 | grep -v "^$" | tail -2
