for sh in 16,0,16,160,192,160 32,16,16,160,192,160 64,32,32,80,96,80; do
 for fl in 0 2; do
  echo "== flags $fl"
  DA_UMMA_FLAGS=$fl DA_UMMA_DEBUG=1 DA_SHAPE=$sh python tools/profile_conv.py
 done
 DA_CONV_SPLIT=tf32 DA_UMMA_DEBUG=1 DA_SHAPE=$sh python tools/profile_conv.py
done
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "test_conv3d" 2>&1 | tail -3
