# generic P1-tet operators (plain fused kernel with node copies, 80-row blocks) against the build before the persistent kernel
export AB_REPS=12
L=$PWD/fdapde-core_b200/lib
run() { echo "== $*"; env "$@" timeout 150 python tools/ab_assembly.py 2>&1 | grep -E "median|rror|Traceback" | cut -c1-200; }
for op in adr diff mass; do
run AB_CONFIG=c4 AB_OP=$op
run AB_CONFIG=c4 AB_OP=$op FDB_LIB_PATH=$L/libfdapde_b200_prepersist.so
done
