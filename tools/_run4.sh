timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "test_conv3d" 2>&1 | tail -8
for sh in 16,0,16,160,192,160 32,16,16,160,192,160 64,32,32,80,96,80 32,0,32,80,96,80 64,64,64,40,48,40 8,0,16,160,192,160; do
  DA_NO_DGRAD=1 DA_SHAPE=$sh timeout 120 python tools/profile_conv.py
  DA_NO_DGRAD=1 DA_CONV_SPLIT=tf32 DA_SHAPE=$sh timeout 120 python tools/profile_conv.py
done
