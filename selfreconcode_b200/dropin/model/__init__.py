from .network import ImplicitNetwork, getTmpSdf
from .Deformer import CompositeDeformer, MLPTranslator, LBSkinner, getTranslatorNet
from .RenderNet import RenderingNetwork_view_norm, getRenderNet
from .CameraMine import RectifiedPerspectiveCameras
from .optim import OptimNetwork
