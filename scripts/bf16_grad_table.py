"""Per-tensor bf16 gradient error table at the benchmarked clip size (nt = 256) against the real reference's fp32 gradients
(tests/golden/full_nt256_b{2,8}.npz).  Run on the GPU box; writes gpurun_out/bf16_grad_table.txt.

    python scripts/bf16_grad_table.py [fixture ...]
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("step_tests", os.path.join(ROOT, "tests", "test_step_gpu.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)

import torch  # noqa: E402

out = []
for fx in (sys.argv[1:] or ["full_nt256_b8", "full_nt256_b2"]):
    auto, auto_loss = T.autocast_gradient_errors(fx)
    ae = sorted(auto.values())
    out.append(f"# {fx} oracle under torch.autocast(bfloat16) on this GPU (the reference's own mixed-precision route): loss {auto_loss:.6f}; "
               f"median {ae[len(ae) // 2]:.2e}  p90 {ae[int(len(ae) * 0.9)]:.2e}  worst {ae[-1]:.2e}  over 2e-2: {sum(e >= 2e-2 for e in ae)}")
    for dtype in (torch.bfloat16, torch.float32):
        rows, loss, ref_loss, m = T.bf16_gradient_table(fx, dtype)
        errs = sorted(r[1] for r in rows)
        out.append(f"# {fx} {dtype}: loss {loss:.6f} (reference {ref_loss:.6f}, rel {abs(loss - ref_loss) / ref_loss:.2e}); gradient tensors {len(rows)}: "
                   f"median {errs[len(errs) // 2]:.2e}  p90 {errs[int(len(errs) * 0.9)]:.2e}  worst {errs[-1]:.2e}  over 2e-2: {sum(e >= 2e-2 for e in errs)}")
        for k, e, n, ne in sorted(rows, key=lambda r: -r[1])[: (40 if dtype == torch.bfloat16 else 8)]:
            out.append(f"  {e:9.2e}  (autocast {auto[k]:9.2e})  norm-err {ne:9.2e}  |g| {n:9.3e}  {k}")
        del m
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "bf16_grad_table.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
