# persistent fused kernel as default: tests + bench + ncu capture.   gpurun --timeout 1200 -- 'bash tools/r2_call9.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or bit_identical or operators or reproducible" > gpurun_out/r2e_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/r2e_quick.log
timeout 400 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python profiles/pick.py < gpurun_out/r2e_bench.json
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_gputests.log
export AB_REPS=15 FDB_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "fused plan: rb|persistent|median|rror|Traceback" ; }
run AB_CONFIG=c4
run AB_CONFIG=c4 FDB_FUSED_RB=80 FDB_FUSED_SMEM_KB=88
run AB_CONFIG=c4 FDB_FUSED_RB=76 FDB_FUSED_SMEM_KB=88
run AB_CONFIG=c4 FDB_FUSED_RB=68
run AB_CONFIG=c4 FDB_PERSIST_NT=320
run AB_CONFIG=c4 AB_OP=mass
run AB_CONFIG=c4 AB_OP=mass FDB_FUSED_PERSIST=0
NCU="ncu --set full --clock-control none --import-source on -f"
name=r02_ncu_persist_c4
AB_REPS=2 AB_CONFIG=c4 timeout 300 $NCU -k regex:k_fused_persist -s 4 -c 1 -o gpurun_out/$name python tools/ab_assembly.py > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep; echo "ncu done"
