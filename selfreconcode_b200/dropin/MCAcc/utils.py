"""Grid helpers (reference: MCAcc/utils.py:88-101,133-146)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def create_grid3D(min, max, steps, device="cuda:0"):
    if type(min) is int:
        min = (min, min, min)
    if type(max) is int:
        max = (max, max, max)
    if type(steps) is int:
        steps = (steps, steps, steps)
    ax = [torch.linspace(min[a], max[a], steps[a]).long().to(device) for a in range(3)]
    gridD, gridH, gridW = torch.meshgrid([ax[2], ax[1], ax[0]], indexing="ij")
    return torch.stack([gridW, gridH, gridD]).view(3, -1).t()  # [N,3] (x,y,z)


class SmoothConv3D(nn.Module):
    """3-D box filter (kept for API parity; Seg3dLossless no longer uses a conv to dilate)."""

    def __init__(self, in_channels, out_channels, kernel_size=3):
        super().__init__()
        assert kernel_size % 2 == 1, "kernel_size for smooth_conv must be odd: {3, 5, ...}"
        self.padding = (kernel_size - 1) // 2
        self.register_buffer('weight', torch.ones((in_channels, out_channels, kernel_size,
                                                   kernel_size, kernel_size),
                                                  dtype=torch.float32) / (kernel_size ** 3))

    def forward(self, input):
        return F.conv3d(input, self.weight, padding=self.padding)
