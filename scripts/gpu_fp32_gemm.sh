python -m pytest tests/test_gpu_mlp.py tests/test_gpu_render.py -x -q -m gpu -k "fp32 or golden or ragged or train_step or generic" 2>&1 | tail -3
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from nerfail_b200 import _lib, ops
lib = _lib.load()
dev = torch.device('cuda')
for (M, N, K) in [(786432, 256, 256), (786432, 256, 319), (786432, 256, 63)]:
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev); Y = torch.empty(M, N, device=dev)
    f = lambda: lib.nfb_linear_fwd(X.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), M, N, K, 1, Y.data_ptr(), N, ops.stream())
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ref = torch.relu(X[:1000] @ W.t() + b)
    print(f"fwd M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  maxerr {float((Y[:1000]-ref).abs().max()):.2e}")
PY
