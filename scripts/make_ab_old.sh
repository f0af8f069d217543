#!/bin/bash
# extract the committed HEAD into ab_old/ and build it there, for the in-call A/B scripts (gpu_ab.sh, gpu_ab2.sh);
# ab_old/ is scratch: delete it after the measurement (it is not tracked)
set -e
cd "$(dirname "$0")/.."
rm -rf ab_old && mkdir ab_old
git archive HEAD | tar -x -C ab_old
(cd ab_old && python -m deepsolid_b200.build | tail -1)
[ -f MEASURED_PEAKS.json ] && cp MEASURED_PEAKS.json ab_old/ || true
