"""char_recognition/model.rs mirror: Net::new(&vs.root()) + ModuleT::forward_t
(model.rs:12-39), with the softmax(-1, Double)/topk(1) of mod.rs:53-56 available fused."""
import ctypes as C

import numpy as np

from .. import _ffi


class Net:
    def __init__(self, weights, ctx=None):
        self.ctx = ctx if ctx is not None else _ffi.default_context()
        n, names, data, numel, keep = _ffi.weights_to_c(weights)
        self._h = _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_rec_create(self.ctx.handle, n, names, data, numel, C.byref(self._h)))

    def forward_t(self, xs, train=False):
        """xs float32 [B,784] in [0,1] -> logits float32 [B,62]."""
        if train:
            raise NotImplementedError("training is out of the hot path")
        return self.predict(xs, want=("logits",))[0]

    def predict(self, xs, want=("logits", "argmax", "prob")):
        """-> (logits [B,62] f32, argmax [B] i32, prob [B] f64); xs float32 [B,784] or uint8 [B,784]."""
        xs = np.ascontiguousarray(xs)
        B = xs.size // 784
        logits = np.empty((B, 62), np.float32) if "logits" in want else None
        argmax = np.empty(B, np.int32) if "argmax" in want else None
        prob = np.empty(B, np.float64) if "prob" in want else None
        L = _ffi.lib()
        if xs.dtype == np.uint8:
            _ffi.check(L.ocrb_rec_forward_u8(self._h, _ffi.ptr(xs), B, _ffi.ptr(logits), _ffi.ptr(argmax), _ffi.ptr(prob)))
        else:
            xs = np.ascontiguousarray(xs, np.float32)
            _ffi.check(L.ocrb_rec_forward(self._h, _ffi.ptr(xs), B, _ffi.ptr(logits), _ffi.ptr(argmax), _ffi.ptr(prob)))
        return logits, argmax, prob

    def close(self):
        if self._h:
            _ffi.lib().ocrb_rec_destroy(self._h)
            self._h = _ffi.c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
