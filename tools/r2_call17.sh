# bank-aware cell positions (FDB_FUSED_BANKS=1) under the persistent kernel.   gpurun --timeout 600 -- 'bash tools/r2_call17.sh'
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 150 python tools/ab_assembly.py 2>&1 | grep -E "persistent|median|rror|Traceback" | cut -c1-220; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_FUSED_BANKS=1
run AB_CONFIG=c2
run AB_CONFIG=c2 FDB_FUSED_BANKS=1
