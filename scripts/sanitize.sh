#!/bin/bash
# compute-sanitizer passes over the smoke step (fp32 + bf16/tcgen05 forward / backward, front-end, loss) and the small front-end / model tests.
# Run on the GPU box: bash scripts/sanitize.sh ; summaries land in gpurun_out/sanitizer_*.txt (copy to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  echo "== $tool smoke: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_smoke.log | tail -1)  | $(grep -c 'smoke ok' gpurun_out/sanitizer_${tool}_smoke.log) smoke ok"
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_frontend_gpu.py tests/test_frozen_gpu.py -m gpu -q -x -k "not full_batch" > gpurun_out/sanitizer_memcheck_tests.log 2>&1
echo "== memcheck frontend+frozen tests: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_tests.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_memcheck_tests.log | tail -1)"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x -k "not full_batch and not 262400" > gpurun_out/sanitizer_racecheck_frontend.log 2>&1
echo "== racecheck frontend tests: $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck_frontend.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_racecheck_frontend.log | tail -1)"
