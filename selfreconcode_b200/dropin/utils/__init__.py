from .FindSurfacePs import *
from .utils import *
