# split (x, y) / z node copies: tests + A/B.   gpurun --timeout 1200 -- 'bash tools/r2_call10.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or bit_identical or operators or reproducible" > gpurun_out/r2f_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/r2f_quick.log
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "fused plan: rb|persistent|median|rror|Traceback" ; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_PERSIST_NT=320
run AB_CONFIG=c4 FDB_FUSED_RB=76 FDB_FUSED_SMEM_KB=88
run AB_CONFIG=c4 FDB_FUSED_RB=80 FDB_FUSED_SMEM_KB=88
run AB_CONFIG=c4 FDB_FUSED_RB=80 FDB_FUSED_SMEM_KB=88 FDB_PERSIST_NT=320
run AB_CONFIG=c4 FDB_FUSED_RB=80 FDB_FUSED_SMEM_KB=88 FDB_PERSIST_NT=448
run AB_CONFIG=c2
run AB_CONFIG=c2 FDB_FUSED_NODES=0
timeout 400 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python profiles/pick.py < gpurun_out/r2f_bench.json
