#!/bin/bash
# GPU parity suite, one pytest process per file so that a device fault in one file cannot hide the others' reports.
OUT=gpurun_out/${1:-tests}
mkdir -p $OUT
rc=0
for f in tests/test_gpu_*.py; do
    b=$(basename $f .py)
    timeout 600 python -m pytest $f -q --timeout 180 -rf > $OUT/$b.log 2>&1; r=$?
    echo "$b exit $r" | tee -a $OUT/summary.log
    [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|E  )|Error|exit" $OUT/$b.log | head -40; }
done
exit $rc
