"""coordx_b200 - B200-native implementation of the CoordX split-MLP hot path.

`coordx_b200.modules` mirrors the reference's `modules.py` API; the arithmetic runs in
`libcoordx_b200.so` (include/coordx_b200.h)."""
from . import _lib  # noqa: F401
from .modules import SingleBVPNet, SplitFCBlock, BatchLinear, Sine, PosEncodingNeRF  # noqa: F401

__version__ = '0.1.0'
