"""text_detection/mod.rs mirror (inference part)."""
import numpy as np

from . import metrics, model  # noqa: F401
from .metrics import get_boxes_and_box_scores
from .model import resnet18

MODEL_FILENAME = "text_detection.model"  # text_detection/mod.rs:18
DEFAULT_WIDTH = 800
DEFAULT_HEIGHT = 800


def run_text_detection(rgba_image, weights, dimensions=(DEFAULT_WIDTH, DEFAULT_HEIGHT), mode="bf16", ctx=None):
    """run_text_detection (text_detection/mod.rs:23-69) minus file IO: decoded RGBA image and
    a name->array weight dict instead of paths.  Returns (pred [1,1,h,w], PolygonScores)."""
    from .. import image_ops
    pre, adj_x, adj_y = image_ops.preprocess_image(rgba_image, dimensions, ctx)
    net = resnet18(weights, mode=mode, ctx=ctx)
    w, h = dimensions
    pred = net.forward_t(pre.reshape(1, 1, h, w), False)
    res = get_boxes_and_box_scores(pred, np.array([[adj_x, adj_y]], np.float64), ctx=ctx)
    return pred, res
