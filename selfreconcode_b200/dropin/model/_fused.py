"""Glue between nn.Module parameters and the fused engine (ops.FusedMLP)."""
import torch

from selfreconcode_b200 import ops
from selfreconcode_b200._lib import SR_ACT_NONE, SR_ACT_SOFTPLUS100, SR_ACT_RELU, SR_ACT_TANH


def ratio_value(ratio, key):
    if isinstance(ratio, dict):
        return ratio[key]
    return ratio


def needs_autograd(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: CUDA tensor required -- selfrecon_b200 has no CPU path" % what)


class FoldCache:
    """Re-folds (weight-norm, transpose, pad) when any parameter changed in place or was
    replaced; optimiser steps bump Tensor._version, load_state_dict copies in place too."""

    def __init__(self):
        self.sig = None
        self.net = None
        self.extra = {}

    def get(self, params, builder):
        sig = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        if sig != self.sig:
            same_storage = self.sig is not None and self.net is not None and \
                tuple((a[0], a[2]) for a in sig) == tuple((a[0], a[2]) for a in self.sig)
            if same_storage and hasattr(self.net, "refold"):
                # only the VALUES changed (optimizer.step(), load_state_dict): refold into the same device buffers
                # -- descriptors, tensor-core packs, truncated views and captured graphs stay valid
                self.net.refold()
            else:
                self.net = builder()
                self.extra = {}
            self.sig = sig
        return self.net
