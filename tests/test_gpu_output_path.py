"""Sampler output path: the fused grid + uint8 kernel against torchvision's make_grid / save_image,
which is what the reference samplers call (ddpm_inference.py:41-45, ddpim_inference.py:90-93)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("N,C,H,W,nrow,pad", [(1, 3, 16, 16, 8, 2), (5, 3, 16, 24, 2, 2), (16, 3, 64, 64, 4, 2),
                                              (36, 3, 32, 32, 6, 2), (7, 1, 8, 8, 3, 1), (9, 3, 16, 16, 16, 0),
                                              (256, 3, 64, 64, 16, 2)])
def test_image_grid_matches_torchvision(N, C, H, W, nrow, pad):
    import torchvision.utils as vutils
    from ddpm_diffusion_model_b200.testing._common import image_grid
    torch.manual_seed(N * 7 + H)
    x = torch.rand(N, C, H, W, device=_dev()) * 1.2 - 0.1          # a few values outside [0,1]: the uint8 clamp matters
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 0.5, 127.5 / 255.0], device=_dev())[:min(4, W)]
    grid, host = image_grid(x, nrow, pad)
    ref = vutils.make_grid(x, nrow=nrow, padding=pad)
    assert grid.shape == ref.shape and torch.equal(grid, ref)
    ref_u8 = ref.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8)       # save_image's conversion
    assert torch.equal(host, ref_u8)


def test_saved_files_decode_to_the_reference_pixels(tmp_path):
    import torchvision.utils as vutils
    from PIL import Image
    from ddpm_diffusion_model_b200.testing._common import flush_image_writes, save_each, save_grid
    torch.manual_seed(3)
    x = torch.rand(10, 3, 32, 32, device=_dev())
    g = save_grid(x, 4, str(tmp_path / "ours.png"))
    vutils.save_image(vutils.make_grid(x, nrow=4, padding=2), str(tmp_path / "ref.png"))
    a, b = np.asarray(Image.open(tmp_path / "ours.png")), np.asarray(Image.open(tmp_path / "ref.png"))
    assert a.shape == b.shape and (a == b).all()
    assert torch.equal(g, vutils.make_grid(x, nrow=4, padding=2))
    save_each(x, str(tmp_path / "each"))
    for i in (0, 9):
        vutils.save_image(x[i], str(tmp_path / f"r{i}.png"))
        assert (np.asarray(Image.open(tmp_path / "each" / f"img_{i:03d}.png")) == np.asarray(Image.open(tmp_path / f"r{i}.png"))).all()
    # asynchronous writer: same bytes once flushed
    os.environ["DDPM_B200_ASYNC_IO"] = "1"
    try:
        save_grid(x, 4, str(tmp_path / "async.png"))
        flush_image_writes()
    finally:
        os.environ.pop("DDPM_B200_ASYNC_IO")
    assert (tmp_path / "async.png").read_bytes() == (tmp_path / "ours.png").read_bytes()


def test_cpu_tensors_are_refused():
    from ddpm_diffusion_model_b200.testing._common import image_grid
    with pytest.raises(RuntimeError):
        image_grid(torch.rand(4, 3, 8, 8), 2)
