"""lgteun_b200 — B200-native (sm_100a) forward of LGTEUN's stage-wise unfolding network.

Public surface (mirrors the reference's hot-path interface, models/unlg_former.py):
    Pansharpening(cfg, logger, stage)   drop-in nn.Module, forward(ms, pan) -> HrMS
    install(...)                        plug it behind the reference's MODELS registry / UnlgFormer runner
    shard_range / forward_sharded       batch sharding for one-process-per-GPU inference
    forward_scene / plan_tiles          overlap-tile driver for scenes larger than one tile
    Trainer / FlatParameters            the training step (train-mode forward, L1, backward, all-reduce, Adam)
    Freprocess(channels)                companion operator SFIIN.Freprocess (FFT amplitude / phase fusion), forward(msf, panf)
    WindowAttention(...)                companion operator: PanFormer's (shifted, cross) window attention, forward(x, y=None)
    swap_companions(net)                replace those two operators inside an unmodified reference network (SFIIN, PanFormer)
The compute lives in lgteun_b200/csrc (CUDA) behind the C ABI of include/lgteun.h."""
from . import _abi
from .module import Pansharpening, expected_state_dict_keys, param_count
from .register import install
from .hostio import HostPipeline
from .sharding import forward_sharded, shard_range
from .scene import forward_scene, plan_tiles
from .train import FlatParameters, Trainer
from .companions import Freprocess, WindowAttention, swap_companions

__all__ = ["Pansharpening", "install", "shard_range", "forward_sharded", "forward_scene", "plan_tiles", "HostPipeline", "Trainer", "FlatParameters", "Freprocess", "WindowAttention", "swap_companions", "expected_state_dict_keys", "param_count", "_abi"]
__version__ = "0.1.0"
