#!/bin/bash
# Config 1 of BASELINE.json: the zultra CLI, zlib format, on the 48 KB minified-JS-shaped input; our CLI against the reference's
python - <<'PY'
import sys; sys.path.insert(0, '.')
from zultra_b200 import synth
open('/tmp/js48k.bin', 'wb').write(synth.js48k().tobytes())
PY
t() { local a=$(date +%s.%N); "$@" > /dev/null 2>&1; local b=$(date +%s.%N); python -c "print('%.3f s' % ($b - $a))"; }
for i in 1 2 3; do echo "b200 cli (process start, CUDA context, compress, write): $(t zultra_b200/zultra -z -zlib /tmp/js48k.bin /tmp/js48k.b200.zz)"; done
if [ -x oracle/_ref/zultra_ref ]; then for i in 1 2; do echo "reference cli: $(t oracle/_ref/zultra_ref -z -zlib /tmp/js48k.bin /tmp/js48k.ref.zz)"; done; cmp /tmp/js48k.b200.zz /tmp/js48k.ref.zz && echo "byte-identical: $(stat -c %s /tmp/js48k.ref.zz) bytes"; fi
python - <<'PY'
import sys, time; sys.path.insert(0, '.')
import zultra_b200 as z
from zultra_b200 import synth
d = synth.js48k()
z.memory_compress(d, 1)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); o = z.memory_compress(d, 1); ts.append(time.perf_counter() - t0)
print("warm in-process zultra_memory_compress of the same 48 944 B: %.2f ms (median of 20), %d bytes out" % (sorted(ts)[10] * 1e3, len(o)))
PY
