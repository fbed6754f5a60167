# Full GPU test suite followed by the one-GPU evidence (superset of capture_n1.sh, with CLI traces).
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest28.log 2>&1
tail -4 gpurun_out/r2_pytest28.log
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:'rs_scatter_k|rs_hist_k|lcp_pack' -c 5 -o /tmp/ncu_a python tools/ncu_one.py enwik100m > /dev/null 2>&1
$NCU -k regex:'unit_dist_k|zb_mf_scan_k|zb_mf_text_k' -c 4 -o /tmp/ncu_b python tools/ncu_one.py enwik100m > /dev/null 2>&1
$NCU -k regex:'zb_parse_dp_k|zb_cand_k|zb_sweep_k' -c 4 -o /tmp/ncu_c python tools/ncu_one.py enwik100m > /dev/null 2>&1
for r in a b c; do python tools/ncu_summary.py /tmp/ncu_$r.ncu-rep > gpurun_out/r2_ncu28_$r.summary.txt 2>&1; done
python tools/ncu_lines.py /tmp/ncu_c.ncu-rep zb_parse_dp_k 40 > gpurun_out/r2_ncu28_parse_dp.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep zb_mf_scan_k 40 > gpurun_out/r2_ncu28_mf_scan.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep zb_mf_text_k 25 > gpurun_out/r2_ncu28_mf_text.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep unit_dist_k 25 > gpurun_out/r2_ncu28_unit_dist.lines.txt 2>&1
python tools/ncu_traffic.py gpurun_out/r2_ncu28_traffic.json /tmp/ncu_a.ncu-rep /tmp/ncu_b.ncu-rep /tmp/ncu_c.ncu-rep > /dev/null 2>&1
cat gpurun_out/r2_ncu28_traffic.json
python -c "
import json
j = json.load(open('gpurun_out/r2_ncu28_traffic.json'))
assert j.get('parse_dp', 0) > 0
json.dump(j, open('profiles/ncu_traffic.json', 'w'), indent=1, sort_keys=True)
"
( time timeout 900 python bench.py ) > gpurun_out/r2_bench28_n1.json 2> gpurun_out/r2_bench28_n1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench28_ref.json 2> gpurun_out/r2_bench28_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches28_enwik100m.csv python bench.py --steps 2 --warmup 1 --strong "" --no-cpu-baseline > gpurun_out/r2_ncu28_bench.log 2>&1
python -c "
import sys; sys.path.insert(0,'.')
from zultra_b200 import synth
import bench
open('/tmp/js.bin','wb').write(synth.js48k().tobytes())
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
"
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -zlib /tmp/js.bin /tmp/js.z ) ; done; ( time ./oracle/_ref/zultra_ref -zlib /tmp/js.bin /tmp/js_ref.z ); cmp /tmp/js.z /tmp/js_ref.z && echo same ) > gpurun_out/r2_cli28_js48k.txt 2>&1
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done; ls -l /tmp/enwik.gz ) > gpurun_out/r2_cli28_enwik.txt 2>&1
timeout 900 python tools/bench_batch.py --count 100000 --steps 2 --out gpurun_out/r2_batch100k_28.jsonl > gpurun_out/r2_batch100k_28.log 2>&1
tail -2 gpurun_out/r2_batch100k_28.log | cut -c1-600
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe28.jsonl > gpurun_out/r2_probe28.log 2>&1
du -sh gpurun_out
