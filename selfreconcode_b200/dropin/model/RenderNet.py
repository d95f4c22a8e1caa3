"""IDR-style rendering network (reference: model/RenderNet.py:9-95) on the fused engine.

Keeps the reference's constructor, attribute names and state_dict keys.  Without autograd the
289->512x4->3 stack (PE of the view direction, ReLU, tanh) runs as one fused kernel
(csrc/mlp_kernels.cu: render_kernel); with autograd it runs as differentiable torch ops."""
import torch
import torch.nn as nn

from selfreconcode_b200 import ops
from .Embedder import get_embedder
from ._fused import (FoldCache, needs_autograd, ratio_value, require_cuda, SR_ACT_NONE, SR_ACT_RELU,
                     SR_ACT_TANH)


class RenderingNetwork_view_norm(nn.Module):
    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True,
                 multires_n=0, multires_v=0):
        super().__init__()
        self.mode = mode
        dims = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.embedv_fn = None
        self.multires_v = multires_v
        if multires_v > 0:
            self.embedv_fn, input_ch = get_embedder(multires_v)
            dims[0] += (input_ch - 3)
        self.embedn_fn = None
        self.multires_n = multires_n
        if multires_n > 0:
            self.embedn_fn, input_ch = get_embedder(multires_n)
            dims[0] += (input_ch - 3)
        self.num_layers = len(dims)
        self.weight_norm = weight_norm
        self.d_in0 = dims[0]
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.relu = nn.ReLU()
        self.tanh = nn.Tanh()
        self._cache = FoldCache()

    def fused(self, ratio):
        if self.mode != 'idr' or self.multires_n > 0 or self.multires_v <= 0:
            raise RuntimeError("RenderingNetwork_view_norm: the fused engine implements mode='idr' "
                               "with multires_v>0, multires_n=0 (config.conf defaults)")
        layers = []
        for l in range(self.num_layers - 1):
            lin = getattr(self, "lin" + str(l))
            v, g = (lin.weight_v, lin.weight_g) if self.weight_norm else (lin.weight, None)
            layers.append(dict(v=v, g=g, b=lin.bias,
                               act=SR_ACT_RELU if l < self.num_layers - 2 else SR_ACT_TANH,
                               skip=False))
        params = [t for L in layers for t in (L["v"], L["g"], L["b"]) if t is not None]
        require_cuda(params[0], "RenderingNetwork_view_norm")

        def build():
            return ops.FusedMLP(self.d_in0, self.multires_v, params[0].device).fold(layers)

        net = self._cache.get(params, build)
        net.set_pe_weights(ops.annealing_weights(self.multires_v, ratio_value(ratio, "renderRatio")))
        return net

    def _train_ok(self):
        from selfreconcode_b200 import train_ops
        return train_ops.TC_TRAIN_ENABLED and self._fusable()

    def forward_train(self, points, normals, view_dirs, feature_vectors, ratio):
        """Differentiable colours on the tensor-core training engine (inputs and parameters)."""
        from selfreconcode_b200 import train_ops as T
        L = self.num_layers - 1
        lins = [getattr(self, "lin" + str(l)) for l in range(L)]
        Ws = T.weight_norm_all(lins) if self.weight_norm else [lin.weight for lin in lins]
        bs = [lin.bias for lin in lins]
        pe_w = ops.annealing_weights(self.multires_v, ratio_value(ratio, "renderRatio"))
        ev = T.embed_rows(view_dirs, self.multires_v, pe_w, 1, ld=3 + 6 * self.multires_v)
        x = torch.cat([points, ev, normals, feature_vectors], dim=-1)
        x0 = torch.nn.functional.pad(x, (0, (-x.shape[1]) % 32))
        acts = [SR_ACT_RELU] * (L - 1) + [SR_ACT_NONE]
        packs = ops.tc_net(self.fused(ratio)).layers
        out = T.tc_mlp(x0, T.MlpConfig(acts, [False] * L, x.shape[1], 1, packs), Ws, bs)
        return torch.tanh(out)

    def _fusable(self):
        """Configurations the fused engines implement (config.conf's defaults): mode 'idr', PE on the view direction
        only, layers at most 512 wide.  Anything else runs the torch ops below (same math, any configuration)."""
        wide = max(getattr(self, "lin" + str(l)).bias.shape[0] for l in range(self.num_layers - 1))
        return self.mode == 'idr' and self.multires_n == 0 and self.multires_v > 0 and wide <= 512

    def forward(self, points, normals, view_dirs, feature_vectors, ratio):
        require_cuda(points, "RenderingNetwork_view_norm.forward")
        if self._train_ok() and needs_autograd(points, normals, view_dirs, feature_vectors, *self.parameters()):
            return self.forward_train(points, normals, view_dirs, feature_vectors, ratio)
        if self._fusable() and not needs_autograd(points, normals, view_dirs, feature_vectors, *self.parameters()):
            return ops.render_forward(self.fused(ratio), points, normals, view_dirs, feature_vectors)
        ratio = ratio_value(ratio, 'renderRatio')

        def emb(fn, x, L):
            if fn is None:
                return x
            if ratio is None:
                return fn(x)
            if ratio <= 0:
                return fn(x, [0.0] * (L * 2))
            return fn(x, [w for w in ops.annealing_weights(L, ratio) for _ in (0, 1)])

        view_dirs = emb(self.embedv_fn, view_dirs, self.multires_v)
        normals = emb(self.embedn_fn, normals, self.multires_n)
        if self.mode == 'idr':
            x = torch.cat([points, view_dirs, normals, feature_vectors], dim=-1)
        elif self.mode == 'no_view_dir':
            x = torch.cat([points, normals, feature_vectors], dim=-1)
        elif self.mode == 'no_normal':
            x = torch.cat([points, view_dirs, feature_vectors], dim=-1)
        for l in range(self.num_layers - 1):
            x = getattr(self, "lin" + str(l))(x)
            if l < self.num_layers - 2:
                x = self.relu(x)
        return self.tanh(x)


def getRenderNet(device, conf):
    return RenderingNetwork_view_norm(conf.get_int('condlen'), d_in=9, d_out=3, dims=[512] * 4,
                                      mode='idr', weight_norm=True,
                                      multires_v=conf.get_int('multires_v'),
                                      multires_n=conf.get_int('multires_n')).to(device)
