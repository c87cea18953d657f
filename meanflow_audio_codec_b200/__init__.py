"""B200-native (sm_100a) implementation of the meanflow_audio_codec hot path.

MDCT analysis -> iMF training step on MDCT tokens -> few-NFE sampling -> IMDCT
overlap-add, as hand-written CUDA behind the reference's Python signatures and a
C ABI (``include/mfac.h``, ``libmfac.so``).  See DESIGN.md.
"""
from ._lib import LIB_PATH, MfacError  # noqa: F401
from .mdct import MDCTConfig, MDCTLayer, IMDCTLayer, imdct, mdct  # noqa: F401
