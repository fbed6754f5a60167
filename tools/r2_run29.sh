set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --strong-steps 2 ) > gpurun_out/r2_bench29_n2.json 2> gpurun_out/r2_bench29_n2.err
tail -c 400 gpurun_out/r2_bench29_n2.json; tail -3 gpurun_out/r2_bench29_n2.err
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2_pytest29_multi.log 2>&1
tail -3 gpurun_out/r2_pytest29_multi.log
