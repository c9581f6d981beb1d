# Development helper: sweep of the fused-plan tunables (rows per block, threads per CTA).
#   gpurun --timeout 900 -- 'bash tools/sweep_fused.sh'
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "fused plan|median|rror" ; }
for c in c2 c3; do
run AB_CONFIG=$c
run AB_CONFIG=$c FDB_FUSED_SMEM_KB=72
run AB_CONFIG=$c FDB_FUSED_SMEM_KB=72 FDB_FUSED_THREADS=384
run AB_CONFIG=$c FDB_FUSED_SMEM_KB=100 FDB_FUSED_THREADS=512
run AB_CONFIG=$c FDB_FUSED_SMEM_KB=32
done
run AB_CONFIG=c4 FDB_FUSED_RB=64 FDB_FUSED_SMEM_KB=72 FDB_FUSED_THREADS=352
run AB_CONFIG=c4 FDB_FUSED_RB=64 FDB_FUSED_SMEM_KB=72 FDB_FUSED_THREADS=416
run AB_CONFIG=c4 FDB_FUSED_RB=64 FDB_FUSED_SMEM_KB=72 FDB_FUSED_THREADS=320
