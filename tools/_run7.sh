for sh in 16,0,16,160,192,160 32,0,16,160,192,160; do
 for fl in 0 4 8 12 1; do
  echo "== flags $fl"
  DA_UMMA_FLAGS=$fl DA_UMMA_DEBUG=1 DA_SHAPE=$sh DA_FWD_ONLY=1 python tools/profile_conv.py 2>&1 | grep -v "^  wgrad"
 done
done
timeout 900 python -m pytest tests/test_gpu_nets.py -q -m gpu 2>&1 | tail -8
