# A/B of the block-local node copies at 3 CTAs per SM.   gpurun --timeout 900 -- 'bash tools/r2_call6.sh'
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "fused plan: (thr|rb)|median|rror|Traceback" ; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=60
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=56
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=54
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=52
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=48
run AB_CONFIG=c4 FDB_FUSED_RB=56
run AB_CONFIG=c2 FDB_FUSED_NODES=1 FDB_FUSED_SMEM_KB=36
run AB_CONFIG=c2 AB_OP=mass FDB_FUSED_NODES=1
run AB_CONFIG=c2 AB_OP=mass
