# compute-sanitizer memcheck (all small paths) and racecheck (48 KB stream): gpurun --timeout 1800 -- 'bash tools/capture_sanitizer.sh'
set -x
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py all ) > gpurun_out/r2_sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitize_memcheck.txt
tail -5 gpurun_out/r2_sanitize_memcheck.txt
( time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_small.py js ) > gpurun_out/r2_sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitize_racecheck.txt
tail -5 gpurun_out/r2_sanitize_racecheck.txt
