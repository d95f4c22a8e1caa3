"""Positional encoding with per-band annealing weights (reference: model/Embedder.py:4-55).

Only used on the autograd (training) path; the inference / tracing path evaluates the encoding
inside the fused kernels (selfreconcode_b200/csrc/mlp_core.cuh: embed_point)."""
import torch


def get_embedder(multires):
    freqs = [float(2.0 ** b) for b in range(multires)]
    out_dim = 3 + 6 * multires

    def embed(x, ws=None):
        parts = [x]
        i = 0
        for f in freqs:
            for fn in (torch.sin, torch.cos):
                v = fn(x * f)
                if ws is not None:
                    v = ws[i] * v
                parts.append(v)
                i += 1
        return torch.cat(parts, -1)

    return embed, out_dim
