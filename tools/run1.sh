export AB_REPS=25
L=$PWD/fdapde-core_b200/lib
run() { echo "== $*"; env "$@" timeout 200 python tools/ab_assembly.py 2>&1 | grep -E "median|rror|Trace" ; }
run AB_CONFIG=c2
run AB_CONFIG=c2 FDB_LIB_PATH=$L/libfdapde_b200_old.so
run AB_CONFIG=c2 AB_OP=mass
run AB_CONFIG=c2 AB_OP=mass FDB_LIB_PATH=$L/libfdapde_b200_old.so
