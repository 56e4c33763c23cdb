for v in w1 w2 w8; do
  echo "variant=$v"
  DFPSR_LIB=$PWD/dfpsr_b200/variants/libdfpsr_b200_$v.so DFPSR_ASYNC=1 python tools/tile_ab.py 256 2>&1 | grep -E "batch|single"
done
