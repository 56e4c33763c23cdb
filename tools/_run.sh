for v in d7 d9 d10; do echo "== $v"; DFPSR_LIB=dfpsr_b200/variants/libdfpsr_b200_$v.so python tools/tile_ab.py 256 2>&1 | grep "exact\] batch"; done
