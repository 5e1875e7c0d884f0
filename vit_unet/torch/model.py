"""Drop-in module path of the reference (``import vit_unet.torch.model as models``, run_denoising.py:2,78).

Everything here is executed by the B200 CUDA kernels of ``vit_unet_b200``; see INTEGRATION.md.
"""
from vit_unet_b200.model import HViT_UNet, ViT_UNet, get_vit_unet  # noqa: F401
