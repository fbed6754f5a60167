set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest38.log 2>&1
tail -3 gpurun_out/r2_pytest38.log
timeout 300 python bench.py --steps 5 --warmup 3 --strong "" --no-cpu-baseline > gpurun_out/r2_bench38_n1.json 2> gpurun_out/r2_bench38_n1.err
tail -c 200 gpurun_out/r2_bench38_n1.json
