for sh in 16,0,16,160,192,160 32,16,16,160,192,160 64,32,32,80,96,80; do
  DA_UMMA_DEBUG=1 DA_NO_DGRAD=1 DA_SHAPE=$sh timeout 120 python tools/profile_conv.py
done
