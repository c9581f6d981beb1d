L=$PWD/fdapde-core_b200/lib
for v in "" _nopipe; do
  FDB_LIB_PATH=$L/libfdapde_b200$v.so timeout 300 python tools/ab_assembly.py 2>&1 | tail -1
done
for t in 192 256; do
  echo "threads $t"; FDB_FUSED_THREADS=$t timeout 300 python tools/ab_assembly.py 2>&1 | tail -1
done
