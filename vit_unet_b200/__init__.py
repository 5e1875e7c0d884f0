"""vit_unet_b200 -- B200-native (sm_100a) forward/backward of the ViT-UNet model of benayas1/vit-unet.

Public surface (mirrors ``vit_unet.torch.model`` of the reference):
    ViT_UNet, HViT_UNet, get_vit_unet      nn.Modules executed by hand-written CUDA kernels
    l1_loss, mse_loss, dice_loss           fused losses
    set_precision('fp32' | 'tf32' | 'bf16')   CUDA-core exact path / tcgen05 TF32 path / bf16 storage + bf16 tcgen05 products
"""
from .engine import get_precision, set_bf16_maps, set_bf16_probs, set_map_l2_budget, set_precision, set_streamed
from .losses import DiceLoss, L1Loss, MSELoss, dice_loss, l1_loss, mse_loss
from .model import HViT_UNet, ViT_UNet, get_vit_unet
from .input_pipeline import DenoisingBatchPipeline
from .optim import FusedAdamW

__all__ = ["ViT_UNet", "HViT_UNet", "get_vit_unet", "l1_loss", "mse_loss", "dice_loss", "L1Loss", "MSELoss",
           "DiceLoss", "set_precision", "get_precision", "FusedAdamW"]
