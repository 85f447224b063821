"""Stage-by-stage comparison of the fp32-mode GPU render against the CPU oracle on the golden 12x12 view."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NERFAIL_B200_MLP"] = "fp32"
import nerfail_b200 as nb
from nerfail_b200 import ops
from oracle import nerf_oracle as no, synth

dev = torch.device("cuda:0")
g = dict(np.load("tests/golden/render.npz"))
H, W = int(g["H"]), int(g["W"])
sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
nets = []
for sd in (sd_c, sd_f):
    n = nb.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).to(dev)
    n.load_state_dict(sd); nets.append(n)
e10, _ = nb.get_embedder(10); e4, _ = nb.get_embedder(4)
q = nb.NetworkQuery(e10, e4, 1 << 16)

def rep(name, a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    d = (a - b).abs()
    print(f"{name:14s} max abs {float(d.max()):.3e}  rel-to-max {float(d.max() / (b.abs().max() + 1e-30)):.3e}  mean abs {float(d.mean()):.3e}")

with torch.no_grad():
    rays_o = no.camera_rays(H, W, g["K"], torch.tensor(g["c2w"]), 2.0, 6.0)
    rays = ops.get_ray_batch(H, W, g["K"], torch.tensor(g["c2w"]), 2.0, 6.0, device=dev)
    rep("rays", rays, rays_o)
    o, d, v = rays_o[:, 0:3], rays_o[:, 3:6], rays_o[:, 8:11]
    z_o = no.coarse_depths(rays_o, 64)
    raw_o = no.query_network(sd_c, o[:, None] + d[:, None] * z_o[..., None], v)
    rgb0_o, disp0_o, acc0_o, w0_o, _ = no.composite(raw_o, z_o, d, True)
    zf_o, zs_o, zstd_o = no.hierarchical_depths(z_o, w0_o, 128)
    rawf_o = no.query_network(sd_f, o[:, None] + d[:, None] * zf_o[..., None], v)
    rgb_o, disp_o, acc_o, w_o, depth_o = no.composite(rawf_o, zf_o, d, True)
    z = ops.coarse_z(rays, 64); rep("z_coarse", z, z_o)
    raw = q.from_rays(rays, z, nets[0]); rep("raw_coarse", raw, raw_o)
    raw_b = q.from_rays(rays_o.to(dev), z_o.to(dev), nets[0]); rep("raw_coarse|orc", raw_b, raw_o)
    rgb0, disp0, acc0, w0, _ = ops.composite_fwd(raw, z, rays, None, True)
    rep("w_coarse", w0, w0_o); rep("rgb0", rgb0, rgb0_o); rep("disp0", disp0, disp0_o)
    c_b = ops.composite_fwd(raw_o.to(dev), z_o.to(dev), rays_o.to(dev), None, True); rep("w_coarse|orc", c_b[3], w0_o)
    zf, zs, zstd = ops.hierarchical(z, w0, 128); rep("z_samples", zs, zs_o); rep("z_fine", zf, zf_o)
    h_b = ops.hierarchical(z_o.to(dev), w0_o.to(dev), 128); rep("z_samples|orc", h_b[1], zs_o); rep("z_fine|orc", h_b[0], zf_o)
    worst = (h_b[1].cpu() - zs_o).abs().max(dim=1)
    r = int(worst.values.argmax()); k = int((h_b[1].cpu()[r] - zs_o[r]).abs().argmax())
    print("worst sample: ray", r, "k", k, "gpu", float(h_b[1][r, k]), "oracle", float(zs_o[r, k]))
    rawf = q.from_rays(rays, zf, nets[1]); rep("raw_fine", rawf, rawf_o)
    rawf_b = q.from_rays(rays_o.to(dev), zf_o.to(dev), nets[1]); rep("raw_fine|orc", rawf_b, rawf_o)
    rgb, disp, acc, w, depth = ops.composite_fwd(rawf, zf, rays, None, True)
    rep("rgb", rgb, rgb_o); rep("disp", disp, disp_o); rep("acc", acc, acc_o); rep("depth", depth, depth_o)
    f_b = ops.composite_fwd(rawf_o.to(dev), zf_o.to(dev), rays_o.to(dev), None, True)
    rep("rgb|orc", f_b[0], rgb_o); rep("disp|orc", f_b[1], disp_o)
    pts = (o[:, None] + d[:, None] * zf_o[..., None]).reshape(-1, 3)
    rep("embed10", e10(pts.to(dev)), no.positional_encoding(pts, 10))
