#!/bin/bash
# One GPU session: FP64 peak micro-benchmark, GPU parity tests, smoke, short bench.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 ./profiles/fp64_peak > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err
cat gpurun_out/fp64_peak.json
timeout 300 python - > gpurun_out/dgemm.json 2>&1 <<'PY'
import torch, json, time
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps(dict(dgemm_8192_tflops=2 * n ** 3 / (best * 1e-3) / 1e12, ms=best)))
PY
cat gpurun_out/dgemm.json
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
