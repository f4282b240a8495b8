"""modarith_b200 -- B200-native batched finite-field and Montgomery-ladder engine behind the
generated-code API of mcarrickscott/modarith.  See DESIGN.md.

    from modarith_b200 import Field, x25519, x448

Importing the package does not touch the GPU; `Field(...)` / `x25519(...)` load the CUDA
library and fail loudly if it is missing (there is no CPU fallback)."""
import importlib

__all__ = ["Field", "x25519", "x448"]


def __getattr__(name):
    if name == "Field":
        return importlib.import_module(".field", __name__).Field
    if name in ("x25519", "x448"):
        return getattr(importlib.import_module(".rfc7748", __name__), name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
