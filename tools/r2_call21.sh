# persistent kernel on the general reference-tensor rows of P1 tetrahedra (ADR, diffusion tensor)
export AB_REPS=12
run() { echo "== $*"; env "$@" timeout 150 python tools/ab_assembly.py 2>&1 | grep -E "median|rror|Traceback" | cut -c1-200; }
for op in adr diff; do
run AB_CONFIG=c4 AB_OP=$op
run AB_CONFIG=c4 AB_OP=$op FDB_FUSED_PERSIST=0
done
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "bit_identical or operators or varying" 2>&1 | tail -2
