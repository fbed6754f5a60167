# bench.py on 8 GPUs under torchrun (weak headline + mix1g / mozilla51m strong) and the multi-GPU tests.
# gpurun --gpus 8 --timeout 1500 -- 'bash tools/capture_n8.sh'
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --strong-steps 2 ) > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
tail -c 300 gpurun_out/scale_n8.json; tail -3 gpurun_out/scale_n8.err
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi_n8.log 2>&1
tail -3 gpurun_out/pytest_multi_n8.log
