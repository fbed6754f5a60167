set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2_pytest6_multi8.log 2>&1
tail -4 gpurun_out/r2_pytest6_multi8.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 2 --strong-steps 2 ) > gpurun_out/r2_bench6_n8.json 2> gpurun_out/r2_bench6_n8.err
tail -c 800 gpurun_out/r2_bench6_n8.err
