# quick confirmation of a build: the parity tests that finish in a minute, the stage probe, the smoke test
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest_quick.log 2>&1
tail -3 gpurun_out/r2_pytest_quick.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m batch10k --out gpurun_out/r2_probe_quick.jsonl > gpurun_out/r2_probe_quick.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_quick.log 2>&1; tail -2 gpurun_out/r2_smoke_quick.log
