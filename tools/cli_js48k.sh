#!/bin/bash
# Config 1 of BASELINE.json: the zultra CLI, zlib format, on the 48 KB minified-JS-shaped input; our CLI against the reference's
python - <<'PY'
import sys; sys.path.insert(0, '.')
from zultra_b200 import synth
open('/tmp/js48k.bin', 'wb').write(synth.js48k().tobytes())
PY
for i in 1 2 3; do /usr/bin/time -f "b200 cli %e s" zultra_b200/zultra -z -zlib /tmp/js48k.bin /tmp/js48k.b200.zz 2>&1 | tail -1; done
if [ -x oracle/_ref/zultra_ref ]; then for i in 1 2; do /usr/bin/time -f "reference cli %e s" oracle/_ref/zultra_ref -z -zlib /tmp/js48k.bin /tmp/js48k.ref.zz 2>&1 | tail -1; done; cmp /tmp/js48k.b200.zz /tmp/js48k.ref.zz && echo "byte-identical: $(stat -c %s /tmp/js48k.ref.zz) bytes"; fi
