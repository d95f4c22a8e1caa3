"""Drop-in `utils` package: the hot-path helpers the reference's modules import by `import utils`
(model/network.py:8): ray/surface finder, Jacobian / cardinal-ray / normal helpers, small math."""
from . import FindSurfacePs as _find
from . import utils as _math

for _mod in (_find, _math):
    for _name in dir(_mod):
        if not _name.startswith("_"):
            globals()[_name] = getattr(_mod, _name)
del _mod, _name
