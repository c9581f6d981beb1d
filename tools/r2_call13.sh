# phase-2 unrolling of the persistent kernel; P2 defaults.   gpurun --timeout 900 -- 'bash tools/r2_call13.sh'
export AB_REPS=12 FDB_VERBOSE=1
L=$PWD/fdapde-core_b200/lib
run() { echo "== $*"; env "$@" timeout 150 python tools/ab_assembly.py 2>&1 | grep -E "persistent|median|rror|Traceback" | cut -c1-220; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_LIB_PATH=$L/libfdapde_b200_nounroll.so
run AB_CONFIG=c3
run AB_CONFIG=c3 FDB_LIB_PATH=$L/libfdapde_b200_nounroll.so
run AB_CONFIG=c3 FDB_FUSED_SMEM_KB=32
run AB_CONFIG=c3 FDB_FUSED_SMEM_KB=48
run AB_CONFIG=p2tet FDB_FUSED_PERSIST_P2=1 FDB_FUSED_SMEM_KB=28
run AB_CONFIG=p2tet FDB_FUSED_PERSIST_P2=1 FDB_FUSED_SMEM_KB=22
run AB_CONFIG=p2tet
