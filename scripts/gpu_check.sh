#!/bin/bash
# Runs every GPU test in its own process (a trapped kernel kills the CUDA context of its process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
python - <<'PY' > gpurun_out/collect.txt 2>&1
import subprocess, sys
import os
out = subprocess.run([sys.executable, "-m", "pytest", os.environ.get("GPU_TESTS", "tests"), "-m", "gpu", "--collect-only", "-q"], capture_output=True, text=True).stdout
ids = [l.strip() for l in out.splitlines() if "::" in l]
print("\n".join(ids))
PY
status=0
while read -r id; do
  [ -z "$id" ] && continue
  echo "=== $id" | tee -a gpurun_out/tests.log
  timeout 300 python -m pytest "$id" -x -q 2>&1 | tail -n 25 | tee -a gpurun_out/tests.log
  rc=${PIPESTATUS[0]}
  [ $rc -ne 0 ] && status=1
done < gpurun_out/collect.txt
echo "overall status $status" | tee -a gpurun_out/tests.log
exit $status
