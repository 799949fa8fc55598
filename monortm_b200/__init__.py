"""monortm_b200 -- B200 (sm_100a) implementation of monoRTM's optical-depth + radiance hot path.

The product is libmonortm_b200.so (hand-written CUDA behind the C ABI of include/monortm_b200.h);
this package is the thin host side: ctypes bindings, a mirror of the reference's MODM / CALCTMR /
RTM operator interface, the TAPE3 / MONORTM_PROF.IN harness readers and the synthetic inputs.
There is no CPU fallback: importing works anywhere, computing needs the built library and a B200.
"""
from . import _capi, linefile, synth, profio, sharding  # noqa: F401
from .api import MonortmError, Session, scor_for_layers, tips_2003  # noqa: F401

__version__ = "0.1.0"
