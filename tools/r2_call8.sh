# persistent fused kernel: CTA sizes + ncu capture.   gpurun --timeout 900 -- 'bash tools/r2_call8.sh'
export AB_REPS=15 FDB_VERBOSE=1 FDB_FUSED_NODES=1 FDB_FUSED_PERSIST=1
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_assembly.py 2>&1 | grep -E "persistent|median|rror|Traceback" ; }
run AB_CONFIG=c4 FDB_PERSIST_NT=384
run AB_CONFIG=c4 FDB_PERSIST_NT=448
run AB_CONFIG=c4 FDB_PERSIST_NT=640 FDB_FUSED_RB=96 FDB_FUSED_SMEM_KB=104
run AB_CONFIG=c4 FDB_PERSIST_NT=384 FDB_FUSED_RB=56
run AB_CONFIG=c4 FDB_PERSIST_NT=384 FDB_FUSED_RB=72 FDB_FUSED_SMEM_KB=80
NCU="ncu --set full --clock-control none --import-source on -f"
name=r02_ncu_persist_c4
AB_REPS=2 AB_CONFIG=c4 timeout 300 $NCU -k regex:k_fused_persist -s 4 -c 1 -o gpurun_out/$name python tools/ab_assembly.py > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep; echo "ncu done"
