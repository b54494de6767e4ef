"""text_detection/model.rs mirror: resnet18(&vs.root()) -> FuncT with forward_t
(model.rs:65-156).  The VarStore is a dict of float32 arrays under the reference's
variable names; the graph itself lives in libocrb (detector.cu)."""
import ctypes as C

import numpy as np

from .. import _ffi


class FuncT:
    def __init__(self, weights, mode="bf16", ctx=None):
        self.ctx = ctx if ctx is not None else _ffi.default_context()
        self.mode = {"fp32": _ffi.MODE_FP32, "bf16": _ffi.MODE_BF16}[mode]
        n, names, data, numel, keep = _ffi.weights_to_c(weights)
        self._h = _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_det_create(self.ctx.handle, n, names, data, numel, self.mode, C.byref(self._h)))

    def forward_t(self, xs, train=False, out=None):
        """xs: [B,1,H,W] (or [B,H,W]) uint8 or float32 raw grey levels, numpy or torch
        (host or cuda) -> probability map float32 [B,1,H,W] of the same kind of container."""
        if train:
            raise NotImplementedError("training is out of the hot path (SURVEY §2 row 2b)")
        shape = tuple(xs.shape)
        if len(shape) == 4:
            if shape[1] != 1:
                raise ValueError("expected [B,1,H,W]")
            B, H, W = shape[0], shape[2], shape[3]
        else:
            B, H, W = shape
        is_torch = hasattr(xs, "data_ptr")
        if is_torch:
            import torch
            dt = {torch.uint8: _ffi.U8, torch.float32: _ffi.F32}[xs.dtype]
            xs = xs.contiguous()
            if out is None:
                out = torch.empty((B, 1, H, W), dtype=torch.float32, device=xs.device)
        else:
            xs = np.ascontiguousarray(xs)
            if xs.dtype not in (np.uint8, np.float32):
                xs = xs.astype(np.float32)
            dt = _ffi.U8 if xs.dtype == np.uint8 else _ffi.F32
            if out is None:
                out = np.empty((B, 1, H, W), np.float32)
        _ffi.check(_ffi.lib().ocrb_det_forward(self._h, _ffi.ptr(xs), dt, B, H, W, _ffi.ptr(out)))
        return out

    def tap(self, name, shape):
        out = np.empty(shape, np.float32)
        _ffi.check(_ffi.lib().ocrb_det_tap(self._h, name.encode(), _ffi.ptr(out), out.size))
        return out

    def close(self):
        if self._h:
            _ffi.lib().ocrb_det_destroy(self._h)
            self._h = _ffi.c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resnet18(weights, mode="bf16", ctx=None):
    """model::resnet18 (model.rs:154-156)."""
    return FuncT(weights, mode, ctx)
