export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 200 python tools/ab_assembly.py 2>&1 | grep -E "fused plan|median|rror|Trace" ; }
run AB_CONFIG=c4
run AB_CONFIG=c2
run AB_CONFIG=c3
run AB_CONFIG=c3 FDB_FUSED_DSM=1
run AB_CONFIG=c3 FDB_FUSED_NOSPLIT=1
run AB_CONFIG=c3 FDB_FUSED_THREADS=192
run AB_CONFIG=p2tet
run AB_CONFIG=p2tet AB_OP=mass
run AB_CONFIG=p2tet FDB_FUSED_NOSPLIT=1
run AB_CONFIG=p2tet FDB_FUSED_THREADS=384
run AB_CONFIG=p2tet FDB_FUSED_SMEM_KB=84 FDB_FUSED_THREADS=256
run AB_CONFIG=p2tet FDB_FUSED_SMEM_KB=84 FDB_FUSED_THREADS=384
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or p2 or operators" 2>&1 | tail -5
