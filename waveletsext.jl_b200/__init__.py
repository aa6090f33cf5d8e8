"""waveletsext.jl_b200 -- B200-native (sm_100a) implementation of the filter-bank hot path of WaveletsExt.jl.

Host mirror of the reference's Julia API for that path (same names; ``f!`` is spelled ``f_``) on top of the C ABI of
libwx_b200.so (include/wx_b200.h).  Import as ``waveletsext_b200`` (the loader module at the repository root) --
the directory name carries a dot and cannot be imported by name.

There is no CPU fallback: importing this package without a built libwx_b200.so raises ImportError, and every
transform requires CUDA tensors.
"""
from . import _lib

_lib.lib()      # fail loudly, at import time, if the CUDA extension has not been built

from .filters import (OrthoFilter, WT, wavelet, makereverseqmfpair, makeqmfpair, autocorr, pfilter, qfilter,  # noqa: E402
                      make_acqmfpair, make_acreverseqmfpair)
from .utils import *          # noqa: E402,F401,F403
from .dwt import *            # noqa: E402,F401,F403
from .swt import *            # noqa: E402,F401,F403
from .acwt import *           # noqa: E402,F401,F403
from .bestbasis import *      # noqa: E402,F401,F403
from .ldb import *            # noqa: E402,F401,F403
from .denoising import *      # noqa: E402,F401,F403
from .siwt import *           # noqa: E402,F401,F403
from . import dist, host      # noqa: E402,F401

launch_count = _lib.launch_count
__version__ = "0.1.0"
