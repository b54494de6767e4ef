"""char_recognition/mod.rs mirror (inference part)."""
from . import model  # noqa: F401
from .model import Net

MODEL_FILENAME = "char_rec_conv_net.model"  # char_recognition/mod.rs:15


def run_prediction(luma_image, weights, ctx=None, k=1):
    """run_prediction (char_recognition/mod.rs:39-68) minus file IO -> (char, probability) for k = 1 (the reference's
    `topk(&output, 1)`), or the k best [(char, probability)] (utils::topk) from the softmax of the logits in f64."""
    import numpy as np

    from .. import image_ops, utils
    net = Net(weights, ctx=ctx)
    x = image_ops.load_image_as_tensor(luma_image, ctx)
    logits, argmax, prob = net.predict(x)
    if k == 1:
        return utils.class_to_char(int(argmax[0])), float(prob[0])
    z = np.asarray(logits, np.float64).reshape(-1, utils.VALUES_COUNT)[0]  # softmax(-1, Kind::Double), mod.rs:53
    e = np.exp(z - z.max())
    return utils.topk(e / e.sum(), k)
