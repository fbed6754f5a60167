set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2_pytest5_multi.log 2>&1
tail -15 gpurun_out/r2_pytest5_multi.log
timeout 300 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe5.jsonl > gpurun_out/r2_probe5.log 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --strong-steps 2 ) > gpurun_out/r2_bench5_n2.json 2> gpurun_out/r2_bench5_n2.err
tail -c 1500 gpurun_out/r2_bench5_n2.err
