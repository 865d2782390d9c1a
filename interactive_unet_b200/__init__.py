"""Import shim: the package lives in `interactive-unet_b200/` (a directory name Python cannot import
directly); this makes it importable as `interactive_unet_b200` with that directory as its path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "interactive-unet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
