"""Training half of the hot path on the tensor-core engine (reference: model/network.py:599-639 losses with
`create_graph=True`, :774-796 parameter VJPs; utils/utils.py:106-120 compute_Jacobian).

The reference obtains grad f and dD/dp by autograd passes THROUGH the networks and then back-propagates a loss
through those passes (double backward).  Here the derivatives w.r.t. the point are carried FORWARD: every point
is four rows of the layer GEMMs -- the value and its three tangents d/dp_x, d/dp_y, d/dp_z -- so f and grad f
(the offset and its Jacobian) are plain outputs of one forward sweep, and every loss built on them needs just ONE
reverse sweep over the same four rows.  That reverse sweep is first order in the rows but second order in the
network (it contains act''), which is exactly what `loss.backward()` over a `create_graph=True` graph computes.

    TcMlpFunction.apply(x0, cfg, W0, b0, W1, b1, ...) -> out
        x0  [M, ld]  fp32 embedded input rows (M = points * ch; ch = 4: value row then three tangent rows; ch = 1)
        W_l [n, k]   effective weights (weight norm already applied by differentiable torch ops outside)
        out [M, n_last]
    forward : pack -> sr_tc_linear per layer (tcgen05 split-bf16, activations kept as tiles)
    backward: pack cotangents -> per layer  sr_tc_wgrad (dW = delta^T x, MN-major tcgen05 GEMM over the kept tiles),
              sr_tc_colsum (db), sr_tc_linear reverse launch (delta_{l-1}; ch = 4 epilogue couples the rows)
No cuBLAS, no torch matmul: torch only carries memory, the embedding and the pointwise loss arithmetic.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import SR_ACT_NONE
from . import ops
from .ops import _p, _pad, _stream, check

INV_SQRT2 = 0.7071067811865476


class MlpConfig:
    """Static description of one MLP stack for TcMlpFunction."""

    def __init__(self, acts, skips, d_in, ch, packs=None):
        self.acts = [int(a) for a in acts]
        self.skips = [bool(s) for s in skips]
        self.d_in = int(d_in)          # width of the embedded input that a skip layer re-appends
        self.ch = int(ch)
        # optional: the module's persistent tensor-core packs (ops.TcNet.layers: W, Wb = W^T, padded bias) of the
        # SAME weights that are passed as tensors -- then nothing is packed per call
        self.packs = packs


def _pack_weights(lib, w):
    w = w.contiguous()
    n, k = w.shape
    buf = torch.empty((lib.sr_tc_weight_bytes(n, k),), dtype=torch.uint8, device=w.device)
    check(lib.sr_tc_pack_weights(_p(w), n, k, w.shape[1], _p(buf), _stream()), "tc_pack_weights")
    return buf


class _Workspace:
    """Scratch shared by backward passes on one (device, stream): wgrad partials, column-sum partials."""
    _pool = {}

    @classmethod
    def get(cls, dev, nbytes):
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
        buf = cls._pool.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
            cls._pool[key] = buf
        return buf


COLSUM_SLICES = 64
DEBUG_LAST = None      # tests: set to a dict to receive the tiles kept by the last forward (acts, shapes)


class TcMlpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, cfg, *wb):
        lib = _lib.load()
        dev = x0.device
        if not x0.is_cuda:
            raise RuntimeError("TcMlpFunction: CUDA tensors required (no CPU path)")
        x0 = x0.detach().contiguous().float()
        M, ld = x0.shape
        L = len(wb) // 2
        Ws = [wb[2 * i].detach().contiguous().float() for i in range(L)]
        bs = [wb[2 * i + 1] for i in range(L)]
        ch = cfg.ch
        with torch.cuda.device(dev):
            A_in = torch.empty((lib.sr_tc_act_bytes(M, ld),), dtype=torch.uint8, device=dev)
            check(lib.sr_tc_pack_rows(_p(x0), M, ld, ld, _p(A_in), None, _stream()), "tc_pack_rows")
            packed, acts, stashes = [], [], []
            cur, K = A_in, ld
            out = None
            for i in range(L):
                n, k = Ws[i].shape
                last = i == L - 1
                pk = cfg.packs[i] if cfg.packs is not None else None
                if pk is not None and (pk["n"], pk["k"]) == (n, k):
                    Wp, bias = pk["W"], pk["bias"]
                else:
                    pk = None
                    Wp = _pack_weights(lib, Ws[i])
                    bias = torch.zeros((_pad(n, 256),), dtype=torch.float32, device=dev)
                    if bs[i] is not None:
                        bias[:n] = bs[i].detach().float()
                packed.append(pk)
                skip_next = (not last) and cfg.skips[i + 1]
                Kn = 0 if last else _pad(Ws[i + 1].shape[1], 32)
                A_next = None if last else torch.empty((lib.sr_tc_act_bytes(M, Kn),), dtype=torch.uint8, device=dev)
                if last:
                    out = torch.empty((M, n), dtype=torch.float32, device=dev)
                # softplus(beta=100): keep act'(z) in fp32 for the reverse sweep (1 - act' recomputed from the
                # 16-bit-mantissa activation tiles would carry 100 x 2^-17 of relative error into every gradient)
                ds = torch.empty((M, _pad(n, 256)), dtype=torch.float32, device=dev) \
                    if (not last and cfg.acts[i] == _lib.SR_ACT_SOFTPLUS100) else None
                check(lib.sr_tc_linear(_p(cur), _p(Wp), _p(bias), M, n, K, n, cfg.acts[i], ch, _p(A_next), Kn,
                                       INV_SQRT2 if skip_next else 1.0, _p(x0) if skip_next else None,
                                       cfg.d_in if skip_next else 0, ld, _p(out), n if last else 0, 0, n, _p(ds), None,
                                       0, 0, 1.0, None, _stream()), "tc_linear")
                if not last:
                    acts.append(A_next)
                    stashes.append(ds)
                    cur, K = A_next, Kn
        ctx.cfg, ctx.M, ctx.ld, ctx.L = cfg, M, ld, L
        ctx.A_in, ctx.acts, ctx.Ws, ctx.packed, ctx.stashes = A_in, acts, Ws, packed, stashes
        ctx.has_bias = [b is not None for b in bs]
        if DEBUG_LAST is not None:
            DEBUG_LAST.update(acts=acts, M=M, widths=[w.shape[0] for w in Ws], kpads=[_pad(w.shape[1], 32) for w in Ws])
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        cfg, M, ld, L, Ws = ctx.cfg, ctx.M, ctx.ld, ctx.L, ctx.Ws
        ch = cfg.ch
        dev = gout.device
        need_x0 = ctx.needs_input_grad[0]
        grads = [None] * (2 * L)
        with torch.cuda.device(dev):
            n_last = Ws[-1].shape[0]
            Kd = _pad(n_last, 32)
            g = gout.detach().contiguous().float()
            D = torch.empty((lib.sr_tc_act_bytes(M, Kd),), dtype=torch.uint8, device=dev)
            check(lib.sr_tc_pack_rows(_p(g), M, n_last, n_last, _p(D), None, _stream()), "tc_pack_rows")
            x0_grad = None
            g_skip = None
            for l in range(L - 1, -1, -1):
                n, k = Ws[l].shape
                X = ctx.acts[l - 1] if l > 0 else ctx.A_in
                Kx = _pad(k, 32) if l > 0 else ld
                # ---- weight / bias gradients from the kept tiles
                if ctx.needs_input_grad[2 + 2 * l]:
                    nbytes = lib.sr_tc_wgrad_partial_bytes(M, Kd, Kx, None)
                    part = _Workspace.get(dev, nbytes)
                    dW = torch.empty((n, k), dtype=torch.float32, device=dev)
                    check(lib.sr_tc_wgrad(_p(D), Kd, _p(X), Kx, M, _p(part), _p(dW), n, k, k, _stream()), "tc_wgrad")
                    grads[2 * l] = dW
                if ctx.has_bias[l] and ctx.needs_input_grad[3 + 2 * l]:
                    ps = torch.empty((COLSUM_SLICES, Kd), dtype=torch.float32, device=dev)
                    check(lib.sr_tc_colsum(_p(D), M, Kd, ch, _p(ps), COLSUM_SLICES, _stream()), "tc_colsum")
                    grads[2 * l + 1] = ps.sum(0)[:n]
                if l == 0 and not need_x0:
                    break
                # ---- reverse GEMM: cotangent of this layer's input
                pk = ctx.packed[l]
                if pk is not None:
                    Wt, zb = pk["Wb"], pk["zero_bias"]
                else:
                    Wt = _pack_weights(lib, Ws[l].t().contiguous())          # [k rows, n cols]
                    zb = torch.zeros((_pad(k, 256),), dtype=torch.float32, device=dev)
                scale = INV_SQRT2 if cfg.skips[l] else 1.0
                if l > 0:
                    n_prev = Ws[l - 1].shape[0]
                    Kd_prev = _pad(n_prev, 32)
                    D_prev = torch.empty((lib.sr_tc_act_bytes(M, Kd_prev),), dtype=torch.uint8, device=dev)
                    if cfg.skips[l]:
                        g_skip = torch.empty((M, _pad(cfg.d_in, 4)), dtype=torch.float32, device=dev)
                    check(lib.sr_tc_linear(_p(D), _p(Wt), _p(zb), M, k, Kd, n_prev, SR_ACT_NONE, ch, _p(D_prev), Kd_prev,
                                           scale, None, 0, 0, _p(g_skip) if cfg.skips[l] else None,
                                           g_skip.shape[1] if cfg.skips[l] else 0, n_prev,
                                           cfg.d_in if cfg.skips[l] else 0, _p(ctx.stashes[l - 1]), _p(ctx.acts[l - 1]),
                                           _pad(k, 32),
                                           cfg.acts[l - 1], scale, None, _stream()), "tc_linear")
                    D, Kd = D_prev, Kd_prev
                else:
                    x0_grad = torch.empty((M, ld), dtype=torch.float32, device=dev)
                    check(lib.sr_tc_linear(_p(D), _p(Wt), _p(zb), M, k, Kd, k, SR_ACT_NONE, 1, None, 0, scale, None, 0,
                                           0, _p(x0_grad), ld, 0, k, None, None, 0, 0, 1.0, None, _stream()),
                          "tc_linear")
                    if k < ld:
                        x0_grad[:, k:] = 0.0
            if x0_grad is not None and g_skip is not None:
                x0_grad[:, :cfg.d_in] += g_skip[:, :cfg.d_in]
        ctx.acts = ctx.A_in = ctx.stashes = None
        return (x0_grad, None) + tuple(grads)


def unpack_tiles(tiles, M, K):
    """tiled split-bf16 activations -> fp32 [M, K] (tests / debugging)."""
    lib = _lib.load()
    out = torch.empty((M, K), dtype=torch.float32, device=tiles.device)
    with torch.cuda.device(tiles.device):
        check(lib.sr_tc_unpack_rows(_p(tiles), M, K, _pad(K, 32), _p(out), K, _stream()), "tc_unpack_rows")
    return out


def tc_mlp(x0, cfg, weights, biases):
    flat = []
    for w, b in zip(weights, biases):
        flat += [w, b]
    return TcMlpFunction.apply(x0, cfg, *flat)


# ------------------------------------------------------------------------------------------------
# Embedding with forward tangents (model/Embedder.py:34-55 + utils/utils.py:40-46), differentiable torch ops
# ------------------------------------------------------------------------------------------------
def embed_rows(p, multires, pe_w, ch, extra=None, ld=None):
    """p [P,3] -> x0 [P*ch, ld]: row 0 of a point = [p, w_k sin(2^k p), w_k cos(2^k p), ..., extra], rows 1..3 =
    d/dp_c of it (zero for `extra`, which does not depend on p).  `extra` [P,E] (latent code, view, ...)."""
    P = p.shape[0]
    freqs = [float(2 ** k) for k in range(multires)]
    vals = [p]
    for k, fr in enumerate(freqs):
        w = float(pe_w[k])
        vals += [w * torch.sin(p * fr), w * torch.cos(p * fr)]
    val = torch.cat(vals, dim=1)                                       # [P, 3+6L]
    pe = val.shape[1]
    E = extra.shape[1] if extra is not None else 0
    width = ld if ld is not None else _pad(pe + E, 32)
    if ch == 1:
        parts = [val] + ([extra] if extra is not None else [])
        row = torch.cat(parts, dim=1)
        return torch.nn.functional.pad(row, (0, width - row.shape[1]))
    eye = torch.eye(3, dtype=p.dtype, device=p.device)
    tans = [eye.unsqueeze(0).expand(P, 3, 3)]                          # d p / d p_c
    for k, fr in enumerate(freqs):
        w = float(pe_w[k]) * fr
        tans += [torch.diag_embed(w * torch.cos(p * fr)), torch.diag_embed(-w * torch.sin(p * fr))]
    tan = torch.cat(tans, dim=2)                                       # [P, 3(c), 3+6L]
    rows = torch.cat([val.unsqueeze(1), tan], dim=1)                   # [P, 4, pe]
    if extra is not None:
        ex = torch.cat([extra.unsqueeze(1), torch.zeros(P, 3, E, dtype=p.dtype, device=p.device)], dim=1)
        rows = torch.cat([rows, ex], dim=2)
    rows = torch.nn.functional.pad(rows, (0, width - rows.shape[2]))
    return rows.reshape(P * 4, width)


def weight_norm_eff(v, g):
    """torch.nn.utils.weight_norm (dim=0): g * v / ||v||_row -- differentiable elementwise ops, no matmul."""
    if g is None:
        return v
    return v * (g.view(-1, 1) / v.norm(dim=1, keepdim=True))


import os as _os

# Fused training path switch (SELFRECON_B200_TC_TRAIN=0 routes training through the torch-autograd twin of the
# same math, which tests use as an A/B reference on the same GPU).
TC_TRAIN_ENABLED = _os.environ.get("SELFRECON_B200_TC_TRAIN", "1") != "0"


def small_matmul(a, b):
    """[...,i,j] x [...,j,k] for tiny trailing dims (3x3, 4x4 bone transforms) as a broadcast multiply + sum:
    stays on elementwise kernels (a torch.matmul here would be a cuBLAS batched GEMM launch per call)."""
    return (a.unsqueeze(-1) * b.unsqueeze(-3)).sum(-2)


def small_matvec(m, v):
    """[...,i,j] x [...,j] -> [...,i] without a GEMM launch."""
    return (m * v.unsqueeze(-2)).sum(-1)


def small_mattvec(m, v):
    """[...,j,i]^T x [...,j] -> [...,i]."""
    return (m * v.unsqueeze(-1)).sum(-2)
