"""homerhevc_b200 -- B200 (sm_100a) implementation of HomerHEVC's ME + interpolation + T/Q hot path.

The product is the C-ABI shared library ``libhomer_b200.so`` (sources in ``csrc/``, header ``include/homer_b200.h``).
This package is only the Python host-side mirror used by the tests and ``bench.py``: thin ctypes bindings whose
names and argument order follow the reference's low-level function table (``low_level_funcs_t``,
/root/reference/src/homer_lib/hmr_private.h:1063) and its encoder-side callers.

There is no CPU fallback: importing works anywhere (so the build and symbol checks can run without a GPU), but every
compute call needs the CUDA library and a visible device and raises otherwise.
"""
from .lib import (HbError, Context, Frame, Prepass, MeJob, MeResult, McJob, McBiJob, TuJob, IntraTuJob, IntraJob, TuResult, TqParams, QuantEnv,
                  PrepassCfg, LowLevelFuncs, FrameIpc, RowSpan, IpcEvent, load_library, library_path, build_library, lowlevel,
                  ME_PEL, ME_HALF, ME_QUARTER, REG_DCT)

__all__ = ["HbError", "Context", "Frame", "Prepass", "MeJob", "MeResult", "McJob", "McBiJob", "TuJob", "IntraTuJob", "IntraJob", "TuResult", "TqParams",
           "QuantEnv", "PrepassCfg", "LowLevelFuncs", "FrameIpc", "RowSpan", "IpcEvent", "load_library", "library_path", "build_library", "lowlevel",
           "ME_PEL", "ME_HALF", "ME_QUARTER", "REG_DCT"]
