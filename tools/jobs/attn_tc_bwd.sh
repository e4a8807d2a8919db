#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" 2>&1 | tail -15 > gpurun_out/r2_tests_attn2.log
tail -15 gpurun_out/r2_tests_attn2.log
if grep -q "passed" gpurun_out/r2_tests_attn2.log && ! grep -q "failed" gpurun_out/r2_tests_attn2.log; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  timeout 200 python - <<'PY' | tee gpurun_out/r2_attn_tc_vs_cuda_core.txt
import sys, torch, ctypes as C
sys.path.insert(0, ".")
from ddpm_diffusion_model_b200 import _lib, engine
dev = torch.device("cuda", 0)
E = engine.Exec(dev, _lib.BF16, False, False)
print("# attention block kernels, bf16, CUDA events over 20 launches (B200): CUDA-core kernels (attn_kernels.cu) vs tcgen05 kernels (attn_tc.cu)")
for (B, H, heads, d) in ((32, 16, 4, 64), (64, 16, 4, 64), (128, 8, 2, 32), (256, 8, 2, 32)):
    inner = heads * d; N = H * H
    qkv = E.act(B, H, H, 3 * inner); qkv.interior().normal_()
    do = E.act(B, H, H, inner); do.interior().normal_()
    o = E.act(B, H, H, inner); lse = E.f32(B, heads, N); dq = E.act(B, H, H, 3 * inner)
    scratch = E.f32(2, B, heads, N, N)
    for simt in (1, 0):
        _lib.lib.ddpm_set_force_simt(simt)
        fw = lambda: _lib.call("ddpm_attn_fwd", C.byref(qkv.desc()), C.byref(o.desc()), heads, d, lse.data_ptr(), _lib.BF16, E.stream)
        bw = lambda: _lib.call("ddpm_attn_bwd", C.byref(qkv.desc()), C.byref(o.desc()), C.byref(do.desc()), lse.data_ptr(), C.byref(dq.desc()), heads, d, scratch.data_ptr(), _lib.BF16, E.stream)
        out = []
        for f in (fw, bw):
            f(); f(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20): f()
            e.record(); torch.cuda.synchronize()
            out.append(s.elapsed_time(e) / 20 * 1e3)
        print(f"B={B:3d} N={N:3d} heads={heads} d={d} {'cuda-core' if simt else 'tcgen05  '}: fwd {out[0]:7.1f} us   bwd {out[1]:7.1f} us")
    _lib.lib.ddpm_set_force_simt(0)
PY
  timeout 100 python bench.py --config celeba256 --steps 8 --warmup 3 --no-eager --no-cpu --no-c256 > gpurun_out/r2_c256_attn_tc.json 2> gpurun_out/r2_c256_attn_tc.err
  python -c "import json;d=json.loads(open('gpurun_out/r2_c256_attn_tc.json').read().strip().splitlines()[-1]);print('c256 train', round(d['value'],1), round(d['ms_per_step'],2), d['loss'])"
fi
