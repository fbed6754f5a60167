#!/bin/bash
# One GPU-box visit: parity tests, bench lines, and (optionally) ncu captures.  usage: gpu_round.sh [tests] [bench] [moz] [ref] [launches] [full]
mkdir -p gpurun_out
for what in "$@"; do case $what in
tests) python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log;;
bench) python bench.py --steps 3 --warmup 3 > gpurun_out/bench_enwik100m.json 2> gpurun_out/bench_enwik100m.err; cut -c1-300 gpurun_out/bench_enwik100m.json; tail -3 gpurun_out/bench_enwik100m.err;;
moz) python bench.py --steps 3 --warmup 3 --workload mozilla51m --no-cpu-baseline > gpurun_out/bench_mozilla51m.json 2> gpurun_out/bench_mozilla51m.err; cut -c1-300 gpurun_out/bench_mozilla51m.json; tail -3 gpurun_out/bench_mozilla51m.err;;
ref) python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_enwik100m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1;;
full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'zb_parse_dp_k' -c 2 -o gpurun_out/full_enwik_dp -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/ncu_full.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'zb_mf_scan_k|zb_mf_text_k' -c 2 -o gpurun_out/full_enwik_mf -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile >> gpurun_out/ncu_full.log 2>&1;;
batch) python tools/bench_batch.py --count 10000 > gpurun_out/batch10k.json 2> gpurun_out/batch10k.err; cut -c1-400 gpurun_out/batch10k.json; tail -2 gpurun_out/batch10k.err;;
scale2) python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; cut -c1-300 gpurun_out/scale_n2.json; tail -3 gpurun_out/scale_n2.err;;
fullmoz) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'zb_parse_fix_k|zb_parse_dp_k' -c 4 -o gpurun_out/full_moz -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --workload mozilla51m > gpurun_out/ncu_fullmoz.log 2>&1;;
esac; done
