#!/bin/bash
# The -m gpu suite, one pytest process per file: a sticky CUDA error (trap / illegal address) in one file cannot cascade into the others.
# usage: bash scripts/gpu_suite_by_file.sh [extra pytest args]
mkdir -p gpurun_out
rc=0
for f in tests/test_*.py; do
  grep -q "pytest.mark.gpu\|mark.gpu" "$f" || continue
  echo "=== $f"; date
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 300 "$@" 2>&1 | tail -40 | grep -vE "^$|Docs: https|warnings summary|UserWarning|Variable._execution" 
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
