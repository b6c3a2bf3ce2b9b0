"""``View`` — drop-in for the reference's common/modules/View.py:4-11."""
import torch.nn as nn

__all__ = ["View"]


class View(nn.Module):
    """Reshape to ``(batch, *target_size)`` in logical NCHW order (layout glue, no kernel)."""

    def __init__(self, *target_size):
        super(View, self).__init__()
        self.target_size = target_size

    def forward(self, input):
        return input.contiguous().view(input.size(0), *self.target_size)

    def extra_repr(self):
        return ", ".join(str(s) for s in self.target_size)
