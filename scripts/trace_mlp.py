"""Timeline of CTA 0 of the fused MLP kernel (nfb_mlp_fwd_trace): where the MMA warp, the producer and the epilogues wait."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfail_b200 import _lib, ops
from oracle import synth

R, S = 65536, 192
dev = torch.device("cuda:0")
m = ops.FusedMLP(device=dev)
m.update(synth.flat_params(synth.make_non_degenerate(synth.random_state_dict(1), 1)).to(dev))
K, _ = synth.intrinsics(800, 800)
rays = ops.get_ray_batch(800, 800, K, torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0, device=dev)[:R].contiguous()
z = ops.coarse_z(rays, S)
raw = torch.empty(R, S, 4, device=dev)
for _ in range(2):
    m.forward_rays(rays, z)
trace = torch.zeros(3, 2048, 4, dtype=torch.int64, device=dev)
_lib.check(_lib.load().nfb_mlp_fwd_trace(m._h, rays.data_ptr(), z.data_ptr(), R, S, raw.data_ptr(), trace.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "trace")
torch.cuda.synchronize()
t = trace.cpu().numpy()
np.save("gpurun_out/mlp_trace.npy", t)
mma = t[1][t[1][:, 1] > 0]
t0 = mma[0, 1]
print("MMA warp events (first 2 units): tag kind/step/idx, wait begin, wait length [cycles]")
tot_ready = tot_w = 0
for e in mma[:120]:
    kind = "A_READY" if (e[0] & 0xF000) == 0x1000 else "W_FULL "
    step, idx = (e[0] >> 4) & 0xFF, e[0] & 0xF
    print(f"  {kind} s={step} i={idx}  t={e[1] - t0:8d}  wait={e[2] - e[1]:6d}")
span = mma[-1, 2] - mma[0, 1]
ready = sum(e[2] - e[1] for e in mma if (e[0] & 0xF000) == 0x1000)
wfull = sum(e[2] - e[1] for e in mma if (e[0] & 0xF000) == 0x2000)
print(f"MMA warp: span {span} cycles over {len(mma)} events; A_READY wait {100 * ready / span:.1f}%  W_FULL wait {100 * wfull / span:.1f}%")
prod = t[0][t[0][:, 1] > 0]
print(f"producer: W_EMPTY wait {100 * sum(e[2] - e[1] for e in prod) / (prod[-1, 2] - prod[0, 1]):.1f}% of its span")
# latency from the producer's issue (end of its event for ring position p) to the MMA warp seeing W_FULL for p
issue = {int(e[0]): int(e[2]) for e in prod}
lat = [int(e[2]) - issue[int(e[3])] for e in mma if (e[0] & 0xF000) == 0x2000 and int(e[3]) in issue and e[2] - e[1] > 50]
if lat:
    print(f"load issue -> W_FULL observed (only when the MMA warp had to wait): median {np.median(lat):.0f} cycles, p90 {np.percentile(lat, 90):.0f}, n={len(lat)}")
epi = t[2][t[2][:, 1] > 0]
acc = [(int(e[0]), int(e[1]), int(e[2])) for e in epi if (e[0] & 0xF000) == 0x3000]
sig = [(int(e[0]), int(e[1])) for e in epi if (e[0] & 0xF000) == 0x4000]
d = []
for (tag, b, en) in acc:
    for (tg, ts) in sig:
        if (tg & 0xFFF) == (tag & 0xFFF) and ts > en:
            d.append(ts - en); break
if d:
    print(f"epilogue duration (ACC_FULL observed -> A_READY signalled): median {np.median(d):.0f} cycles, max {max(d)}")
# per-step epilogue phases of warp 2 (slot 0): ACC_FULL observed -> drain loop done -> bias refilled (2 slot barriers) -> A_READY signalled
ev = [(int(x[0]), int(x[1]), int(x[2])) for x in epi]
rows = []
for i, (tag, b, en) in enumerate(ev):
    if (tag & 0xF000) == 0x3000 and ((tag >> 4) & 0xFF) < 9:
        key = tag & 0xFFF
        t_acc = en; d = {}
        for (tg, b2, e2) in ev[i + 1:i + 5]:
            if (tg & 0xFFF) != key: continue
            if (tg & 0xF800) == 0x6000: d["drain"] = b2 - t_acc
            elif (tg & 0xF800) == 0x6800: d["refill"] = b2 - t_acc
            elif (tg & 0xF000) == 0x4000: d["signal"] = b2 - t_acc
        if len(d) == 3: rows.append((en - b, d["drain"], d["refill"], d["signal"]))
rows = np.array(rows)
if len(rows):
    print("epilogue phases, median cycles after ACC_FULL: drain loop done %d, bias refilled %d, signalled %d; wait for ACC_FULL %d (n=%d)" %
          (np.median(rows[:, 1]), np.median(rows[:, 2]), np.median(rows[:, 3]), np.median(rows[:, 0]), len(rows)))
