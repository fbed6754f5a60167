# bench.py on 2 GPUs under torchrun: gpurun --gpus 2 --timeout 1500 -- 'bash tools/capture_n2.sh'
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --strong-steps 2 ) > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
tail -c 300 gpurun_out/scale_n2.json; tail -3 gpurun_out/scale_n2.err
