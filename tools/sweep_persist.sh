# Development helper: the A/B runs behind the defaults of the persistent fused kernel.  Every line
# times one configuration of tools/ab_assembly.py (L2 flushed between launches) and checks the result bit for bit against
# the two-kernel path.     gpurun --timeout 900 -- 'bash tools/sweep_persist.sh'
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 150 python tools/ab_assembly.py 2>&1 | grep -E "fused plan: rb|persistent|median|rror|Traceback" | cut -c1-220; }
# C4: persistent kernel (default) against the plain fused kernel, block sizes, CTA sizes, bank-aware cell positions
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_FUSED_PERSIST=0
run AB_CONFIG=c4 FDB_FUSED_BANKS=0
for rb in 56 64 72 76; do run AB_CONFIG=c4 FDB_FUSED_RB=$rb; done
for nt in 384 512; do run AB_CONFIG=c4 FDB_PERSIST_NT=$nt; done
# other P1-tetrahedra operators
for op in mass adr diff; do run AB_CONFIG=c4 AB_OP=$op; run AB_CONFIG=c4 AB_OP=$op FDB_FUSED_PERSIST=0; done
# P1 triangles (persistent default, 256-row blocks), P2 triangles (persistent default), P2 tetrahedra (plain default)
run AB_CONFIG=c2; run AB_CONFIG=c2 FDB_FUSED_PERSIST=0; run AB_CONFIG=c2 FDB_FUSED_PERSIST=0 FDB_FUSED_NODES=0; run AB_CONFIG=c2 FDB_FUSED_RB=128; run AB_CONFIG=c2 FDB_FUSED_RB=512 FDB_FUSED_SMEM_KB=100
run AB_CONFIG=c3; run AB_CONFIG=c3 FDB_FUSED_PERSIST_P2=0
run AB_CONFIG=p2tet; run AB_CONFIG=p2tet FDB_FUSED_PERSIST_P2=1
