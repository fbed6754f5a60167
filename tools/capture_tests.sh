set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_final.log 2>&1
tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; tail -2 gpurun_out/r2_smoke_final.log
