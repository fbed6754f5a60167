set -x
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
import bench
bench.gen_workload('enwik100m'); bench.gen_workload('mozilla51m')
"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:'zb_parse_dp_k|zb_mf_scan_k|zb_mf_text_k' -c 6 -o gpurun_out/r2_ncu_enwik python tools/ncu_one.py enwik100m > gpurun_out/r2_ncu_enwik.log 2>&1
$NCU -k regex:'zb_parse_dp_k|zb_mf_scan_k|zb_parse_fix_k' -c 7 -o gpurun_out/r2_ncu_moz python tools/ncu_one.py mozilla51m > gpurun_out/r2_ncu_moz.log 2>&1
ls -la gpurun_out/*.ncu-rep
