AB_SOLVERS=c4 python tools/solver_ab.py 2>&1 | grep -E "C4 CG multi|rel diff persistent" | tail -3
FDB_NO_CG_FUSED_DIR=1 AB_SOLVERS=c4 python tools/solver_ab.py 2>&1 | grep -E "C4 CG multi" | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "solve or cg or bicg or solver or poisson or elliptic or c4 or c2 or c3 or repeat or parabolic or pde" 2>&1 | tail -3
