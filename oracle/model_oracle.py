"""ORACLE (test infrastructure, never on the product path): CPU restatement of the
reference's two networks with the same ATen ops tch 0.3.0 calls.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

Parity status: the reference has no test that pins the network forward
 and the reference binary cannot be built here (no cargo/rustc),
so the forward is "parity unpinned" against ocr-rs itself; what is pinned is that
these are literally the libtorch operators the reference calls, in the reference's
order:

  detector   /root/reference/src/text_detection/model.rs:65-152
  char-rec   /root/reference/src/char_recognition/model.rs:12-39
  softmax    /root/reference/src/char_recognition/mod.rs:53-56, utils.rs:28-43
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5  # tch nn::BatchNormConfig default eps


def _t(w, name):
    v = w[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def _bn(x, w, p):
    # nn::batch_norm2d in eval mode (model.rs:34,42,44,69,102,104)
    return F.batch_norm(x, _t(w, p + ".running_mean"), _t(w, p + ".running_var"),
                        _t(w, p + ".weight"), _t(w, p + ".bias"), False, 0.1, EPS)


def _basic_block(x, w, p, stride, has_down):
    # model.rs:40-55
    y = F.conv2d(x, _t(w, p + ".conv1.weight"), None, stride, 1)
    y = F.relu(_bn(y, w, p + ".bn1"))
    y = F.conv2d(y, _t(w, p + ".conv2.weight"), None, 1, 1)
    y = _bn(y, w, p + ".bn2")
    if has_down:  # model.rs:30-38
        s = F.conv2d(x, _t(w, p + ".downsample.0.weight"), None, stride, 0)
        s = _bn(s, w, p + ".downsample.1")
    else:
        s = x
    return F.relu(y + s)


def _layer(x, w, li, stride):
    x = _basic_block(x, w, f"layer{li}.0", stride, li != 1)
    return _basic_block(x, w, f"layer{li}.1", 1, False)


def _up(x, k):
    n, c, h, wd = x.shape
    return F.interpolate(x, size=(h * k, wd * k), mode="nearest")  # upsample_nearest2d


@torch.no_grad()
def detector_forward(w, x, dtype=torch.float32, return_taps=False):
    """x: [B,1,H,W] float (raw 0..255 grey levels) -> prob [B,1,H,W].

    model.rs:107-151.  `dtype=torch.float64` gives the high-precision reference used to
    measure the oracle's own fp32 rounding noise.
    """
    if dtype != torch.float32:
        w = {k: _t(w, k).to(dtype) for k in w}
    x = torch.as_tensor(x).to(dtype)
    taps = {}
    x1 = F.conv2d(x, _t(w, "conv1.weight"), None, 2, 3)
    x1 = F.relu(_bn(x1, w, "bn1"))
    x1 = F.max_pool2d(x1, 3, 2, 1, 1, False)
    taps["stem"] = x1
    x1 = _layer(x1, w, 1, 1)
    x2 = _layer(x1, w, 2, 2)
    x_in2 = F.conv2d(x1, _t(w, "in2.weight"))
    x3 = _layer(x2, w, 3, 2)
    x_in3 = F.conv2d(x2, _t(w, "in3.weight"))
    x4 = _layer(x3, w, 4, 2)
    x_in4 = F.conv2d(x3, _t(w, "in4.weight"))
    x_in5 = F.conv2d(x4, _t(w, "in5.weight"))
    taps.update(x1=x1, x2=x2, x3=x3, x4=x4)
    # non-cascaded FPN (model.rs:126-138)
    p2 = F.conv2d(_up(x_in3, 2) + x_in2, _t(w, "out2.weight"), None, 1, 1)
    p3 = _up(F.conv2d(_up(x_in4, 2) + x_in3, _t(w, "out3.weight"), None, 1, 1), 2)
    p4 = _up(F.conv2d(_up(x_in5, 2) + x_in4, _t(w, "out4.weight"), None, 1, 1), 4)
    p5 = _up(F.conv2d(x_in5, _t(w, "out5.weight"), None, 1, 1), 8)
    fuse = torch.cat([p5, p4, p3, p2], 1)
    taps["fuse"] = fuse
    y = F.conv2d(fuse, _t(w, "bin_conv1.weight"), None, 1, 1)
    y = F.relu(_bn(y, w, "bin_bn1"))
    taps["bin1"] = y
    y = F.conv_transpose2d(y, _t(w, "bin_conv_tr1.weight"), _t(w, "bin_conv_tr1.bias"), 2, 0)
    y = F.relu(_bn(y, w, "bin_bn2"))
    y = F.conv_transpose2d(y, _t(w, "bin_conv_tr2.weight"), _t(w, "bin_conv_tr2.bias"), 2, 0)
    out = torch.sigmoid(y)
    if return_taps:
        return out, taps
    return out


@torch.no_grad()
def rec_forward(w, x, dtype=torch.float32):
    """x: [B,784] float in [0,1] -> logits [B,62]  (char_recognition/model.rs:27-39;
    eval mode, dropout off; NB no ReLU after the convolutions)."""
    if dtype != torch.float32:
        w = {k: _t(w, k).to(dtype) for k in w}
    x = torch.as_tensor(x).to(dtype).view(-1, 1, 28, 28)
    x = F.conv2d(x, _t(w, "conv1.weight"), _t(w, "conv1.bias"))
    x = F.max_pool2d(x, 2)
    x = F.conv2d(x, _t(w, "conv2.weight"), _t(w, "conv2.bias"))
    x = F.max_pool2d(x, 2)
    x = x.reshape(-1, 1024)
    x = F.relu(F.linear(x, _t(w, "fc1.weight"), _t(w, "fc1.bias")))
    return F.linear(x, _t(w, "fc2.weight"), _t(w, "fc2.bias"))


@torch.no_grad()
def rec_top1(logits):
    """softmax(-1, Double) + topk(1) (char_recognition/mod.rs:53-56, utils.rs:28-43)."""
    p = torch.softmax(logits.to(torch.float64), -1)
    v, i = p.topk(1, -1)
    return i[:, 0].numpy().astype(np.int32), v[:, 0].numpy()


def binarize(pred, thresh=0.6):
    """metrics.rs:129-131: pred.gt(thresh).to_kind(Uint8); the compare is done in f32
    (the f64 scalar is demoted: float32(0.6) > 0.6 is False in libtorch)."""
    return (np.asarray(pred, np.float32) > np.float32(thresh)).astype(np.uint8)
