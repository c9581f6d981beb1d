# L2 persisting window for the Krylov vectors: A/B.   gpurun --timeout 600 -- 'bash tools/r2_call16.sh'
AB_SOLVERS=c4,c3 timeout 250 python tools/solver_ab.py 2>&1 | grep -E "multi-kernel|rel diff"
echo "== FDB_L2_PERSIST=0"
FDB_L2_PERSIST=0 AB_SOLVERS=c4,c3 timeout 250 python tools/solver_ab.py 2>&1 | grep -E "multi-kernel|rel diff"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "solve or cg or bicg or solver or poisson or elliptic or repeat or parabolic or pde" 2>&1 | tail -2
