set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest21.log 2>&1
tail -3 gpurun_out/r2_pytest21.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m batch10k --out gpurun_out/r2_probe21.jsonl > gpurun_out/r2_probe21.log 2>&1
cp zultra_b200/libzultra_b200.so /tmp/main.so
for v in cp512 cp512cg512; do
  cp build/alt/$v.so zultra_b200/libzultra_b200.so
  timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m batch10k --out gpurun_out/r2_probe21_$v.jsonl > gpurun_out/r2_probe21_$v.log 2>&1
done
cp /tmp/main.so zultra_b200/libzultra_b200.so
