"""`src` alias package: put `<repo>/ddpm_diffusion_model_b200/dropin` on PYTHONPATH *instead of* the
reference checkout and code written against the reference (`from src.model.unet_backbone import
UNetDenoiser`, `from src.training_loops.train_one_epoch import train_one_epoch`, ...) runs on the
B200-native implementation unchanged."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

_PKG = "ddpm_diffusion_model_b200"
_MAP = {
    "src.model": f"{_PKG}.model",
    "src.model.difussion_utils": f"{_PKG}.model.difussion_utils",
    "src.model.difussion_class": f"{_PKG}.model.difussion_class",
    "src.model.attention": f"{_PKG}.model.attention",
    "src.model.unet_backbone": f"{_PKG}.model.unet_backbone",
    "src.training_loops": f"{_PKG}.training_loops",
    "src.training_loops.ema": f"{_PKG}.training_loops.ema",
    "src.training_loops.grad_scaler": f"{_PKG}.training_loops.grad_scaler",
    "src.training_loops.training_utils": f"{_PKG}.training_loops.training_utils",
    "src.training_loops.train_one_epoch": f"{_PKG}.training_loops.train_one_epoch",
    "src.training_loops.main_train_loop": f"{_PKG}.training_loops.main_train_loop",
    "src.training_loops.chekpoints": f"{_PKG}.training_loops.chekpoints",
    "src.testing": f"{_PKG}.testing",
    "src.testing.ddpm_inference": f"{_PKG}.testing.ddpm_inference",
    "src.testing.ddpim_inference": f"{_PKG}.testing.ddpim_inference",
}
for _alias, _real in _MAP.items():
    sys.modules[_alias] = importlib.import_module(_real)
