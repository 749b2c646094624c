"""Package shim so that ``import pytorch_prototyping.pytorch_prototyping`` resolves to the drop-in."""
from . import pytorch_prototyping  # noqa: F401

__path__ = []
