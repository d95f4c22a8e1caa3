"""Drop-in `model` package: the names the reference's drivers import (model/__init__.py:1-3:
getTmpSdf, OptimNetwork, getOptNet, initialLBSkinner, RectifiedPerspectiveCameras,
PointsRendererWithFrags) plus the field modules."""
from .network import ImplicitNetwork, getTmpSdf
from .Deformer import (CompositeDeformer, MLPTranslator, LBSkinner, getTranslatorNet, initialLBSkinner,
                       compute_lbswField, smooth_weights)
from .RenderNet import RenderingNetwork_view_norm, getRenderNet
from .CameraMine import RectifiedPerspectiveCameras, PointsRendererWithFrags
from .optim import OptimNetwork, getOptNet
