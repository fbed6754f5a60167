# The evidence kept under profiles/ for one GPU: ncu --set full summaries + source-line tables + DRAM traffic of the hot kernels,
# bench.py (own arm and --impl reference), the ncu launch list of the bench command, the stage probe and the 100 000-payload batch.
# Run on a B200 box: gpurun --timeout 2400 -- 'bash tools/capture_n1.sh'; outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:'rs_scatter_k|rs_hist_k' -c 4 -o /tmp/ncu_a python tools/ncu_one.py enwik100m > /dev/null 2>&1
$NCU -k regex:'unit_dist_k|zb_mf_scan_k|zb_mf_text_k' -c 4 -o /tmp/ncu_b python tools/ncu_one.py enwik100m > /dev/null 2>&1
$NCU -k regex:'zb_parse_dp_k|zb_cand_k|zb_sweep_k' -c 4 -o /tmp/ncu_c python tools/ncu_one.py enwik100m > /dev/null 2>&1
for r in a b c; do python tools/ncu_summary.py /tmp/ncu_$r.ncu-rep > gpurun_out/r2_ncu_$r.summary.txt 2>&1; done
python tools/ncu_lines.py /tmp/ncu_c.ncu-rep zb_parse_dp_k 40 > gpurun_out/r2_ncu_parse_dp.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep zb_mf_scan_k 40 > gpurun_out/r2_ncu_mf_scan.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep zb_mf_text_k 25 > gpurun_out/r2_ncu_mf_text.lines.txt 2>&1
python tools/ncu_lines.py /tmp/ncu_b.ncu-rep unit_dist_k 25 > gpurun_out/r2_ncu_unit_dist.lines.txt 2>&1
python tools/ncu_traffic.py gpurun_out/r2_ncu_traffic.json /tmp/ncu_a.ncu-rep /tmp/ncu_b.ncu-rep /tmp/ncu_c.ncu-rep > /dev/null 2>&1
cat gpurun_out/r2_ncu_traffic.json
python -c "
import json
j = json.load(open('gpurun_out/r2_ncu_traffic.json'))
assert j.get('parse_dp', 0) > 0
json.dump(j, open('profiles/ncu_traffic.json', 'w'), indent=1, sort_keys=True)
"
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_enwik100m.csv python bench.py --steps 2 --warmup 1 --strong "" --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe.jsonl > gpurun_out/r2_probe.log 2>&1
timeout 900 python tools/bench_batch.py --count 100000 --steps 2 --out gpurun_out/r2_batch100k_.jsonl > gpurun_out/r2_batch100k_.log 2>&1
tail -c 500 gpurun_out/r2_bench_n1.json
