#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" 2>&1 | tail -15 > gpurun_out/r2_tests_attn.log
tail -15 gpurun_out/r2_tests_attn.log
if grep -q "passed" gpurun_out/r2_tests_attn.log && ! grep -q "failed" gpurun_out/r2_tests_attn.log; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  timeout 200 python - <<'PY'
import sys, torch, ctypes as C
sys.path.insert(0, ".")
from ddpm_diffusion_model_b200 import _lib, engine
dev = torch.device("cuda", 0)
E = engine.Exec(dev, _lib.BF16, False, False)
for (B, H, heads, d) in ((32, 16, 4, 64), (64, 16, 4, 64), (128, 8, 2, 32), (256, 8, 2, 32)):
    inner = heads * d
    qkv = E.act(B, H, H, 3 * inner); qkv.interior().normal_()
    o = E.act(B, H, H, inner); lse = E.f32(B, heads, H * H)
    for simt in (1, 0):
        _lib.lib.ddpm_set_force_simt(simt)
        f = lambda: _lib.call("ddpm_attn_fwd", C.byref(qkv.desc()), C.byref(o.desc()), heads, d, lse.data_ptr(), _lib.BF16, E.stream)
        f(); f(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): f()
        e.record(); torch.cuda.synchronize()
        print(f"attn fwd B={B} N={H*H} heads={heads} d={d} {'cuda-core' if simt else 'tcgen05 '}: {s.elapsed_time(e)/20*1e3:.1f} us")
    _lib.lib.ddpm_set_force_simt(0)
PY
fi
