"""CPU tests of the host side of the retraining loop (nerfail_b200/train.py, optim.py): the batch sampler, the precrop
window, the learning-rate schedule, checkpoint naming and the Adam state layout.  No kernel is called here — the `-m gpu`
twins (tests/test_gpu_render.py) run the same checks on the device."""
import numpy as np
import torch

from oracle import nerf_oracle as no
from oracle import synth


def test_ray_batch_sampler_on_the_host_matches_reference_sampling():
    """train.sample_ray_batch (device='cpu') against the oracle restatement of run_nerf.py:744-773 under the same numpy
    seed: same image, same pixels (precrop window and full frame), same rays / targets; rank shares partition the batch
    in order, and every rank consumes the host generator identically (so all ranks agree on the image and the pixels)."""
    import nerfail_b200 as nb
    H, W, N = 40, 36, 128
    K, _ = synth.intrinsics(H, W)
    rng = np.random.default_rng(3)
    images = rng.random((6, H, W, 4)).astype(np.float32)
    poses = np.stack(synth.camera_ring(6)).astype(np.float32)
    i_train = [0, 2, 3, 5]
    for step, pre in ((0, 500), (700, 500)):
        ro = np.random.RandomState(11)
        rays_ref, tgt_ref, img_ref, coords_ref = no.sample_ray_batch(images, poses, i_train, H, W, K, N, step, pre, 0.5, rng=ro)
        rg = np.random.RandomState(11)
        rays, tgt, img_i, coords = nb.sample_ray_batch(images, poses, i_train, H, W, K, N, step, pre, 0.5, rng=rg, device="cpu")
        assert img_i == img_ref and torch.equal(coords, coords_ref)
        assert torch.equal(tgt, tgt_ref)
        assert torch.allclose(rays, rays_ref, rtol=0, atol=1e-6)
        assert ro.randint(1 << 30) == rg.randint(1 << 30), "the sampler must consume the generator exactly like the reference"
        if step < pre:
            r0, c0, nr, nc = nb.precrop_window(H, W, 0.5)
            assert int(coords[:, 0].min()) >= r0 and int(coords[:, 0].max()) < r0 + nr
            assert int(coords[:, 1].min()) >= c0 and int(coords[:, 1].max()) < c0 + nc
        parts, states = [], []
        for r in range(3):
            rg = np.random.RandomState(11)
            parts.append(nb.sample_ray_batch(torch.from_numpy(images), poses, i_train, H, W, K, N, step, pre, 0.5,
                                             rng=rg, device="cpu", rank=r, world_size=3))
            states.append(rg.randint(1 << 30))
        assert torch.equal(torch.cat([p[3] for p in parts], 0), coords_ref)
        assert torch.equal(torch.cat([p[1] for p in parts], 0), tgt_ref)
        assert len({p[2] for p in parts}) == 1 and len(set(states)) == 1
        sizes = [p[3].shape[0] for p in parts]
        assert sum(sizes) == N and max(sizes) - min(sizes) <= 1


def test_precrop_window_is_the_reference_centre_crop():
    """run_nerf.py:754-763: dH = int(H//2 * frac), rows H//2-dH .. H//2+dH-1 (2 dH of them), likewise columns."""
    import nerfail_b200 as nb
    for H, W, frac in ((800, 800, 0.5), (100, 100, 0.5), (401, 377, 0.3), (7, 9, 1.0)):
        dH, dW = int(H // 2 * frac), int(W // 2 * frac)
        lin_r = np.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH)
        lin_c = np.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW)
        r0, c0, nr, nc = nb.precrop_window(H, W, frac)
        assert (nr, nc) == (lin_r.size, lin_c.size)
        assert np.array_equal(np.arange(r0, r0 + nr), lin_r.astype(np.int64))
        assert np.array_equal(np.arange(c0, c0 + nc), lin_c.astype(np.int64))


def test_learning_rate_schedule_and_checkpoint_names():
    """run_nerf.py:796-800 (lrate * 0.1 ** (step / (lrate_decay * 1000))) and :809 ('{:06d}.tar')."""
    from nerfail_b200 import optim, train
    for step in (0, 1, 1000, 250000, 500000):
        want = 5e-4 * (0.1 ** (step / (250 * 1000)))
        assert optim.decayed_lrate(5e-4, 250, step) == want
    lin = torch.nn.Linear(3, 2)
    opt = torch.optim.Adam(lin.parameters(), lr=5e-4)
    optim.set_lrate(opt, 1.25e-4)
    assert all(g["lr"] == 1.25e-4 for g in opt.param_groups)
    assert train.checkpoint_path("logs", "lego", 200000).replace("\\", "/") == "logs/lego/200000.tar"
    assert train.checkpoint_path("logs", "lego", 42).endswith("000042.tar")
