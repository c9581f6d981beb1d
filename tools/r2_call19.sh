# 8-GPU C4 line (no extra configs, to stay inside the GPU budget).   gpurun --gpus 8 --timeout 240 -- 'bash tools/r2_call19.sh'
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra > gpurun_out/r2j_bench8.json 2> gpurun_out/r2j_bench8.err; echo "bench8 rc=$?"
grep '^{' gpurun_out/r2j_bench8.json | python profiles/pick.py
tail -3 gpurun_out/r2j_bench8.err
