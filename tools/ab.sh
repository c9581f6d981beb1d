# Development helper: A/B timing of library variants on a GPU box.
#   python fdapde-core_b200/build.py --variant=exp -DSOME_MACRO      (here, then:)
#   gpurun --timeout 300 -- 'bash tools/ab.sh exp'
# Each line reports the fused C4 assembly time of one build and whether it is bit-identical to the two-kernel path.
L=$PWD/fdapde-core_b200/lib
timeout 300 python tools/ab_assembly.py 2>&1 | tail -1
for v in "$@"; do
  FDB_LIB_PATH=$L/libfdapde_b200_$v.so timeout 300 python tools/ab_assembly.py 2>&1 | tail -1
done
