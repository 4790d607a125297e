"""ctypes binding of the C ABI declared in include/lgteun.h (the only way Python talks to the kernels).

The library is built in-tree by ``python -m lgteun_b200.build`` into ``lgteun_b200/_lgteun_cuda.so``.
There is no fallback of any kind: if the library is missing or a call fails, an exception is raised."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lgteun_cuda.so")

RUN_DEAD_PRIORS = 1
NO_GRAPH = 2
EINVAL, ECUDA, ESTATE, ENOMEM = -1, -2, -3, -4

# name -> (restype, argtypes); mirrors include/lgteun.h one to one (checked by tests/test_abi.py)
_F = c_void_p  # float* (device or host)
SIGNATURES = {
    "lgteun_abi_version": (c_int, []),
    "lgteun_last_error": (c_char_p, []),
    "lgteun_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "lgteun_destroy": (None, [c_void_p]),
    "lgteun_num_weights": (c_int, [c_void_p]),
    "lgteun_weight_name": (c_char_p, [c_void_p, c_int]),
    "lgteun_weight_numel": (c_int64, [c_void_p, c_int]),
    "lgteun_load_weights": (c_int, [c_void_p, POINTER(c_char_p), POINTER(c_void_p), POINTER(c_int64), c_int, c_void_p]),
    "lgteun_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int]),
    "lgteun_forward": (c_int, [c_void_p, _F, _F, _F, c_int, c_int, c_int, c_int, c_void_p]),
    "lgteun_forward_host": (c_int, [c_void_p, _F, _F, _F, c_int, c_int, c_int, c_int, c_void_p]),
    "lgteun_forward_launches": (c_int, [c_void_p, c_int, c_int, c_int, c_int]),
    "lgteun_op_bicubic": (c_int, [c_void_p, _F, _F, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_data_step": (c_int, [c_void_p, c_int, _F, _F, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_patch_embed": (c_int, [c_void_p, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_mixer": (c_int, [c_void_p, c_int, c_int, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_local_mixer": (c_int, [c_void_p, c_int, c_int, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_global_mixer": (c_int, [c_void_p, c_int, c_int, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_ffn": (c_int, [c_void_p, c_int, c_int, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_prior": (c_int, [c_void_p, c_int, _F, _F, c_int, c_int, c_int, c_void_p]),
    "lgteun_op_metrics": (c_int, [c_void_p, _F, _F, c_void_p, c_int, c_int, c_int, ctypes.c_float, c_void_p]),
    "lgteun_op_normalize": (c_int, [c_void_p, _F, _F, c_int64, ctypes.c_float, c_void_p]),
    "lgteun_op_to_nhwc": (c_int, [c_void_p, _F, _F, c_int, c_int, c_int, c_int, ctypes.c_float, c_void_p]),
    # training step
    "lgteun_flat_numel": (c_int64, [c_void_p]),
    "lgteun_weight_offset": (c_int64, [c_void_p, c_int]),
    "lgteun_train_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int]),
    "lgteun_train_forward": (c_int, [c_void_p, _F, _F, _F, _F, c_int, c_int, c_int, ctypes.c_float, ctypes.c_uint64, c_void_p]),
    "lgteun_train_backward": (c_int, [c_void_p, _F, _F, c_void_p]),
    "lgteun_train_generation": (ctypes.c_uint64, [c_void_p]),
    "lgteun_train_backward_of": (c_int, [c_void_p, ctypes.c_uint64, _F, _F, c_void_p]),
    "lgteun_train_launches": (c_int, [c_void_p]),
    "lgteun_l1_loss": (c_int, [c_void_p, _F, _F, c_int64, ctypes.c_float, _F, _F, c_void_p]),
    "lgteun_adam_step": (c_int, [c_void_p, _F, _F, _F, _F, c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, c_int, ctypes.c_float, c_void_p]),
    "lgteun_dropout_mask": (c_int, [c_void_p, ctypes.c_uint64, c_int, ctypes.c_float, _F, c_int64, c_void_p]),
    "lgteun_train_set_masks": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lgteun_train_set_seed_ptr": (c_int, [c_void_p, c_void_p]),
    # companion operators
    "lgteun_op_freprocess_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "lgteun_op_freprocess": (c_int, [c_int, _F, _F, _F, c_int, c_int, c_int, c_int, POINTER(c_void_p), _F, c_int64, c_void_p]),
    "lgteun_op_window_attention": (c_int, [c_int, _F, _F, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                           ctypes.c_float, _F, _F, _F, _F, _F, _F, _F, c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m lgteun_b200.build` "
                "(there is no CPU or PyTorch fallback for the LGTEUN forward)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error():
    return (lib().lgteun_last_error() or b"").decode()


def check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc in (EINVAL,):
        raise ValueError(f"lgteun: {msg}")
    if rc == ENOMEM:
        raise MemoryError(f"lgteun: {msg}")
    raise RuntimeError(f"lgteun (code {rc}): {msg}")


class Handle:
    """RAII wrapper of lgteun_t* (one per device, one host thread at a time)."""

    def __init__(self, device, bands, stages):
        self._p = c_void_p()
        self.device, self.bands, self.stages = int(device), int(bands), int(stages)
        check(lib().lgteun_create(self.device, self.bands, self.stages, ctypes.byref(self._p)))

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            lib().lgteun_destroy(self._p)
            self._p = c_void_p()

    __del__ = close

    @property
    def ptr(self):
        return self._p

    def weight_table(self):
        n = lib().lgteun_num_weights(self._p)
        return [(lib().lgteun_weight_name(self._p, i).decode(), lib().lgteun_weight_numel(self._p, i)) for i in range(n)]

    def load_weights(self, tensors, stream=0):
        """tensors: dict name -> contiguous fp32 CUDA tensor on this handle's device."""
        names = list(tensors)
        n = len(names)
        c_names = (c_char_p * n)(*[s.encode() for s in names])
        c_ptrs = (c_void_p * n)(*[tensors[s].data_ptr() for s in names])
        c_nums = (c_int64 * n)(*[tensors[s].numel() for s in names])
        check(lib().lgteun_load_weights(self._p, c_names, c_ptrs, c_nums, n, c_void_p(stream)))

    def forward(self, ms_ptr, pan_ptr, out_ptr, N, h, w, flags=0, stream=0):
        check(lib().lgteun_forward(self._p, ms_ptr, pan_ptr, out_ptr, N, h, w, flags, c_void_p(stream)))

    def forward_host(self, ms_ptr, pan_ptr, out_ptr, N, h, w, flags=0, stream=0):
        check(lib().lgteun_forward_host(self._p, ms_ptr, pan_ptr, out_ptr, N, h, w, flags, c_void_p(stream)))

    def launches(self, N, h, w, flags=0):
        return lib().lgteun_forward_launches(self._p, N, h, w, flags)

    def workspace_bytes(self, N, h, w):
        return lib().lgteun_workspace_bytes(self._p, N, h, w)

    # -- training step -----------------------------------------------------------------------------------------
    def flat_numel(self):
        return lib().lgteun_flat_numel(self._p)

    def flat_layout(self):
        """[(state_dict key, offset in floats, numel)] of the flat parameter / gradient buffers."""
        n = lib().lgteun_num_weights(self._p)
        return [(lib().lgteun_weight_name(self._p, i).decode(), lib().lgteun_weight_offset(self._p, i),
                 lib().lgteun_weight_numel(self._p, i)) for i in range(n)]

    def train_forward(self, flat_param_ptr, ms_ptr, pan_ptr, out_ptr, N, h, w, dropout_p=0.1, seed=0, stream=0):
        check(lib().lgteun_train_forward(self._p, flat_param_ptr, ms_ptr, pan_ptr, out_ptr, N, h, w, dropout_p, seed,
                                         c_void_p(stream)))

    def train_backward(self, dout_ptr, flat_grad_ptr, stream=0, generation=None):
        """Backward of the handle's tape; with ``generation`` (the value of train_generation() right after the forward) it
        raises if another forward has overwritten that tape in the meantime."""
        if generation is None:
            check(lib().lgteun_train_backward(self._p, dout_ptr, flat_grad_ptr, c_void_p(stream)))
        else:
            check(lib().lgteun_train_backward_of(self._p, generation, dout_ptr, flat_grad_ptr, c_void_p(stream)))

    def train_generation(self):
        return lib().lgteun_train_generation(self._p)

    def train_workspace_bytes(self, N, h, w):
        return lib().lgteun_train_workspace_bytes(self._p, N, h, w)

    def train_launches(self):
        return lib().lgteun_train_launches(self._p)

    def l1_loss(self, out_ptr, gt_ptr, n, weight, loss_ptr, dout_ptr, stream=0):
        check(lib().lgteun_l1_loss(self._p, out_ptr, gt_ptr, n, weight, loss_ptr, dout_ptr, c_void_p(stream)))

    def adam_step(self, p_ptr, g_ptr, m_ptr, v_ptr, n, lr, beta1, beta2, eps, step, grad_scale=1.0, stream=0):
        check(lib().lgteun_adam_step(self._p, p_ptr, g_ptr, m_ptr, v_ptr, n, lr, beta1, beta2, eps, step, grad_scale,
                                     c_void_p(stream)))

    def dropout_mask(self, seed, layer, p, out_ptr, n, stream=0):
        check(lib().lgteun_dropout_mask(self._p, seed, layer, p, out_ptr, n, c_void_p(stream)))

    def train_set_seed_ptr(self, seed_dev_ptr):
        check(lib().lgteun_train_set_seed_ptr(self._p, c_void_p(seed_dev_ptr) if seed_dev_ptr else None))

    def set_masks(self, ptrs):
        self.ext_masks = ptrs is not None
        if ptrs is None:
            check(lib().lgteun_train_set_masks(self._p, None))
        else:
            arr = (c_void_p * 5)(*ptrs)
            check(lib().lgteun_train_set_masks(self._p, arr))

    def op(self, name, *args, stream=0):
        check(getattr(lib(), "lgteun_op_" + name)(self._p, *args, c_void_p(stream)))
