timeout 900 python -m pytest tests/test_gpu_nets.py -q -m gpu --tb=line 2>&1 | grep -v "^$" | tail -30
for sh in 16,0,16,160,192,160 32,0,16,160,192,160 32,0,32,80,96,80; do
 for fl in 0 4 8; do
  echo "== flags $fl"
  DA_UMMA_FLAGS=$fl DA_UMMA_DEBUG=1 DA_SHAPE=$sh python tools/profile_conv.py 2>&1
 done
done
