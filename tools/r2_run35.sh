set -x
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 --strong-steps 2 ) > gpurun_out/r2_bench35_n4.json 2> gpurun_out/r2_bench35_n4.err
tail -c 300 gpurun_out/r2_bench35_n4.json; tail -3 gpurun_out/r2_bench35_n4.err
