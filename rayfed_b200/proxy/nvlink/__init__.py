"""NVLink transport: CUDA tensors go GPU→peer-GPU through the symmetric heap."""
from .channel import DeviceChannel  # noqa: F401
