set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest8.log 2>&1
tail -8 gpurun_out/r2_pytest8.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m --out gpurun_out/r2_probe8.jsonl > gpurun_out/r2_probe8.log 2>&1
for v in 1 2 3 4 5 6; do ZULTRA_CUDA_DP_VAR=$v timeout 300 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe8_var$v.jsonl > /dev/null 2>&1; done
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:'zb_mf_scan_k|zb_mf_text_k|tile_filter' -c 6 -o gpurun_out/r2_ncu8_enwik python tools/ncu_one.py enwik100m > gpurun_out/r2_ncu8_enwik.log 2>&1
