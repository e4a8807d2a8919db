"""ddpm_diffusion_model_b200 -- B200-native (sm_100a) implementation of the DDPM/DDIM hot path of
pablo-reyes8/ddpm-diffusion-model behind the reference's own Python API.

    from ddpm_diffusion_model_b200.model.unet_backbone import UNetDenoiser, build_unet_64x64
    from ddpm_diffusion_model_b200.model.difussion_class import Diffusion
    from ddpm_diffusion_model_b200.training_loops.train_one_epoch import train_one_epoch
    from ddpm_diffusion_model_b200.testing.ddpim_inference import ddim_infer_sample

or, as a drop-in for code written against the reference (`from src.model... import ...`), put
`ddpm_diffusion_model_b200/dropin` on PYTHONPATH instead of the reference checkout.

Importing the package loads libddpm_b200.so (building it with nvcc if it is missing) and raises if
that fails: there is no ATen/CPU fallback behind these modules.
"""
from . import _lib  # noqa: F401  (fail loudly at import time)

__all__ = ["_lib"]
__version__ = "0.1.0"
