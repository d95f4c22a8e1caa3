from .network import ImplicitNetwork, getTmpSdf
from .Deformer import CompositeDeformer, MLPTranslator, LBSkinner, getTranslatorNet
from .RenderNet import RenderingNetwork_view_norm, getRenderNet
try:  # orchestration layer (needs pytorch3d for the rasterisers it is handed)
    from .optim import OptimNetwork, getOptNet
except ImportError:  # pragma: no cover
    pass
