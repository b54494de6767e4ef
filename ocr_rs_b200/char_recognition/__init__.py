"""char_recognition/mod.rs mirror (inference part)."""
from . import model  # noqa: F401
from .model import Net

MODEL_FILENAME = "char_rec_conv_net.model"  # char_recognition/mod.rs:15


def run_prediction(luma_image, weights, ctx=None):
    """run_prediction (char_recognition/mod.rs:39-68) minus file IO -> (char, probability)."""
    from .. import image_ops, utils
    net = Net(weights, ctx=ctx)
    x = image_ops.load_image_as_tensor(luma_image, ctx)
    _, argmax, prob = net.predict(x)
    return utils.class_to_char(int(argmax[0])), float(prob[0])
