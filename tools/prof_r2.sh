# Development helper: ncu --set full captures of fused assembly kernels.   gpurun --timeout 900 -- 'bash tools/prof_r2.sh'
NCU="ncu --set full --clock-control none --import-source on -f"
prof() { name=$1; shift; env "$@" AB_REPS=2 FDB_VERBOSE=1 timeout 300 $NCU -k regex:k_fused_assemble -s 4 -c 1 -o gpurun_out/$name python tools/ab_assembly.py 2>&1 | grep -E "fused plan: thr|median|rror"
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null; }
prof r02_p2tet_v4 AB_CONFIG=p2tet
prof r02_c3_v4 AB_CONFIG=c3
