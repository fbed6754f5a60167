set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest15_full.log 2>&1
tail -5 gpurun_out/r2_pytest15_full.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench15_n1.json 2> gpurun_out/r2_bench15_n1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench15_ref.json 2> gpurun_out/r2_bench15_ref.err
timeout 900 python tools/bench_batch.py --count 100000 --steps 2 --out gpurun_out/r2_batch100k.jsonl > gpurun_out/r2_batch100k.log 2>&1
tail -2 gpurun_out/r2_batch100k.log | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_enwik100m.csv python bench.py --steps 2 --warmup 1 --strong "" --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:'zb_parse_dp_k|zb_mf_scan_k|zb_mf_text_k|rs_scatter_k|rs_hist_k|unit_dist_k|lcp_pack' -c 60 -o gpurun_out/r2_ncu15_enwik python tools/ncu_one.py enwik100m > gpurun_out/r2_ncu15_enwik.log 2>&1
python -c "
import sys; sys.path.insert(0,'.')
from zultra_b200 import synth
import bench
open('/tmp/js.bin','wb').write(synth.js48k().tobytes())
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
"
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -zlib /tmp/js.bin /tmp/js.z ) ; done; ( time ./oracle/_ref/zultra_ref -zlib /tmp/js.bin /tmp/js_ref.z ); cmp /tmp/js.z /tmp/js_ref.z && echo same ) > gpurun_out/r2_cli15_js48k.txt 2>&1
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done; ls -l /tmp/enwik.gz ) > gpurun_out/r2_cli15_enwik.txt 2>&1
