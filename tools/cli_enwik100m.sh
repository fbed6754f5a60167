#!/bin/bash
# The zultra CLI (streaming API underneath) on the 100 MB configuration: wall time of the whole process, per staging size
python - <<'PY'
import sys; sys.path.insert(0, '.')
import bench
bench.gen_workload("enwik100m").tofile('/tmp/enwik100m.bin')
PY
t() { local a=$(date +%s.%N); "$@" > /dev/null 2>&1; local b=$(date +%s.%N); python -c "print('%.3f s' % ($b - $a))"; }
zultra_b200/zultra -z /tmp/enwik100m.bin /tmp/e.gz > /dev/null 2>&1   # first run pays the page cache and the CUDA start-up of a cold box
for bb in 64 128 256; do echo "ZULTRA_CUDA_BATCH_BLOCKS=$bb: $(ZULTRA_CUDA_BATCH_BLOCKS=$bb t zultra_b200/zultra -z /tmp/enwik100m.bin /tmp/e$bb.gz) $(stat -c %s /tmp/e$bb.gz) bytes"; done
cmp /tmp/e64.gz /tmp/e256.gz && echo "same bytes for every staging size"
python - <<'PY'
import zlib
assert zlib.decompress(open('/tmp/e64.gz','rb').read(), 31) == open('/tmp/enwik100m.bin','rb').read(); print("gunzip(e64.gz) == input")
PY
