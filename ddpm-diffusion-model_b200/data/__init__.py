"""Input side of the hot path: a device-resident replacement for the reference's CPU loaders."""
from .device_loader import DeviceLoader, dataset_to_u8  # noqa: F401
