"""focal_loss(y_true, y_pred, gamma=2, alpha=0.9) -- host mirror of reference src/utils/focal_loss.py:5-12.
Sum-reduced over the whole batch, clip [1e-3, .999]; one fused CUDA kernel with a deterministic two-stage sum.
Returns a 0-d float64 CUDA tensor."""
from . import ops


def focal_loss(y_true, y_pred, gamma=2, alpha=0.9):
    return ops.focal_loss_sum(y_true.contiguous().float(), y_pred.contiguous().float(), gamma, alpha)[0]
