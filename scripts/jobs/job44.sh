python scripts/variant_time.py
for v in K1 K2 K3 G2 G3 G6; do ACE_B200_LIB=/root/repo/additivecausalexpansion_b200/_variants/libace_$v.so python scripts/variant_time.py; done
