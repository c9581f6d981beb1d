# A/B of the block-local node copies (FDB_FUSED_NODES=1) on C4 / C2.   gpurun --timeout 900 -- 'bash tools/r2_call5.sh'
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "fused plan|median|rror|Traceback" ; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_FUSED_NODES=1
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_THREADS=512
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=96 FDB_FUSED_SMEM_KB=104 FDB_FUSED_THREADS=512
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=48 FDB_FUSED_THREADS=352
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=48 FDB_FUSED_THREADS=320
run AB_CONFIG=c4 FDB_FUSED_NODES=1 FDB_FUSED_RB=32 FDB_FUSED_THREADS=256
run AB_CONFIG=c2
run AB_CONFIG=c2 FDB_FUSED_NODES=1
AB_SOLVERS=c4only timeout 200 python tools/solver_ab.py 2>&1 | grep -E "C4|rel diff"
