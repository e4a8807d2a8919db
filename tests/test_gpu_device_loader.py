"""Input side: the device-resident feeder against the reference's CPU transforms
(ToTensor + Normalize(0.5, 0.5), src/data/load_data_local.py:90-95) and against a DataLoader's shuffle order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_transform(u8_nhwc):
    t = u8_nhwc.permute(0, 3, 1, 2).to(torch.float32).div(255)            # ToTensor
    return (t - 0.5) / 0.5                                                 # Normalize([0.5]*3, [0.5]*3)


def test_batches_are_bit_equal_to_the_reference_transforms_and_sampler_order():
    from torch.utils.data import DataLoader, TensorDataset
    from ddpm_diffusion_model_b200.data import DeviceLoader
    g = torch.Generator().manual_seed(5)
    data = torch.randint(0, 256, (37, 16, 12, 3), dtype=torch.uint8, generator=g)
    data[0, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    ref_ds = TensorDataset(_ref_transform(data), torch.arange(37))
    for drop_last in (False, True):
        ours = DeviceLoader(data, 8, shuffle=True, drop_last=drop_last, device="cuda:0", generator=torch.Generator().manual_seed(11),
                            labels=torch.arange(37), shard=False)
        ref = DataLoader(ref_ds, batch_size=8, shuffle=True, drop_last=drop_last, generator=torch.Generator().manual_seed(11))
        assert len(ours) == len(ref)
        n = 0
        for (xo, yo), (xr, yr) in zip(ours, ref):
            assert xo.is_cuda and xo.dtype == torch.float32 and xo.shape == xr.shape
            assert torch.equal(xo.cpu(), xr) and torch.equal(yo, yr)
            n += 1
        assert n == len(ref)
    # two epochs of the same loader differ (a new permutation per epoch), unshuffled order is the identity
    ours = DeviceLoader(data, 37, shuffle=True, device="cuda:0", generator=torch.Generator().manual_seed(1), shard=False)
    a = next(iter(ours))[0]; b = next(iter(ours))[0]
    assert not torch.equal(a, b)
    plain = DeviceLoader(data, 37, shuffle=False, device="cuda:0", shard=False)
    assert torch.equal(next(iter(plain))[0].cpu(), _ref_transform(data))


def test_dataset_round_trip_and_training_from_the_feeder():
    from ddpm_diffusion_model_b200.data import DeviceLoader, dataset_to_u8
    from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
    from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser
    from ddpm_diffusion_model_b200.training_loops.train_one_epoch import last_step_losses, train_one_epoch
    g = torch.Generator().manual_seed(2)
    u8 = torch.randint(0, 256, (24, 16, 16, 3), dtype=torch.uint8, generator=g)
    floats = [(x, 0) for x in _ref_transform(u8)]                         # what a reference dataset yields
    assert torch.equal(dataset_to_u8(floats), u8)                         # lossless
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = UNetDenoiser(3, 32, (1, 2), 1, {8}, 64, 0.0, 2, 16, 16).to(dev).train()
    d = Diffusion(T=1000).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    loader = DeviceLoader(u8, 8, device="cuda:0", shard=False)
    avg, nb, ni, gs = train_one_epoch(model, d, loader, opt, device="cuda:0")
    assert nb == 3 and ni == 24 and gs == 3 and avg == avg and last_step_losses().numel() == 3


def test_host_tensors_and_cpu_device_are_refused():
    from ddpm_diffusion_model_b200.data import DeviceLoader
    with pytest.raises(RuntimeError):
        DeviceLoader(torch.zeros(4, 8, 8, 3, dtype=torch.uint8), 2, device="cpu")
    with pytest.raises(ValueError):
        DeviceLoader(torch.zeros(4, 8, 8, 3), 2, device="cuda:0")
