"""Times the four tensor-core kernels of a retraining step in isolation (CUDA events, 4096 x 192 samples):
inference forward, training forward (saves images), data-gradient chain, grouped weight gradients."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfail_b200 as nb
from nerfail_b200 import ops, _lib
from oracle import synth

dev = torch.device("cuda:0")
R, S = int(os.environ.get("R", 4096)), int(os.environ.get("S", 192))
net = nb.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).to(dev)
net.load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
K, _ = synth.intrinsics(800, 800)
rays = ops.get_ray_batch(800, 800, K, torch.tensor(synth.camera_ring(8)[1][:3, :4]), 2.0, 6.0, device=dev)[:R].contiguous()
z = torch.linspace(2, 6, S, device=dev).expand(R, S).contiguous()
lib = _lib.load()
fused = net.fused()
M = R * S
T = int(lib.nfb_mlp_train_tiles(M))
act = torch.empty((T, 40, 128, 64), dtype=torch.bfloat16, device=dev)
mask = torch.empty((T, 9, 8, 128), dtype=torch.int32, device=dev)
dy = torch.empty((T, 39, 128, 64), dtype=torch.bfloat16, device=dev)
raw = torch.empty((R, S, 4), device=dev)
g_raw = torch.randn(M, 4, device=dev)
grad = torch.zeros(int(lib.nfb_mlp_param_count(fused._h)), device=dev)
st = ops.stream
P = lambda t: t.data_ptr()
runs = {
    "inference fwd": lambda: lib.nfb_mlp_fwd(fused._h, 1, None, None, P(rays), P(z), R, S, P(raw), st()),
    "train fwd": lambda: lib.nfb_mlp_fwd_train(fused._h, P(rays), P(z), R, S, P(raw), P(act), P(mask), st()),
    "bwd data": lambda: lib.nfb_mlp_bwd_data(fused._h, P(g_raw), M, P(mask), P(dy), st()),
    "bwd weights": lambda: lib.nfb_mlp_bwd_weights(fused._h, P(act), P(dy), P(g_raw), M, P(grad), st()),
    "bwd overlapped": lambda: lib.nfb_mlp_bwd(fused._h, P(g_raw), M, P(mask), P(act), P(dy), P(grad), P(ready), st()),
}
ready = torch.full((T,), 10, dtype=torch.int32, device=dev)
flop = {"inference fwd": 1186816, "train fwd": 1186816, "bwd data": 1115392, "bwd weights": 1186816,
        "bwd overlapped": 1115392 + 1186816}
only = os.environ.get("ONLY")
if only:
    runs = {k: v for k, v in runs.items() if k in only.split(",") or k == "train fwd"}
print(f"skip={os.environ.get('NERFAIL_B200_TRAIN_SKIP', '0')}  producers={os.environ.get('NERFAIL_B200_BWD_PRODUCERS', 'default')}  M={M}")
for name, fn in runs.items():
    for _ in range(2):
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"  {name:14s} {ms:7.3f} ms  {flop[name] * M / ms / 1e9:7.1f} TFLOP/s")
fused.status()
