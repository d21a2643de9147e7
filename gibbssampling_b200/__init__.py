"""gibbssampling_b200 -- B200-native replacement for the hot path of Etschbeijer/GibbsSampling.

Layout (only what the hot path needs):
  csrc/                hand-written sm_100a kernels + the extern "C" boundary (libgibbs_b200.so)
  _abi.py              ctypes binding of include/gibbs_b200.h (what the F# shim binds with P/Invoke)
  engine.py            handle owner: sequences resident in HBM, primitive and chain entry points
  CompositeVector.py   ProbabilityCompositeVector etc.: the types that cross the boundary (fs:11-124)
  SiteSampler.py       reference function names of fs:298-707
  MotifSampler.py      reference function names of fs:709-1038
  distributed.py       chain sharding + the single all_gather of each GPU's best result
  Results.py           what the script does with results: countBy positions, segments, PWM / information content
"""
from . import _abi
from ._abi import (GibbsArgumentError, GibbsCudaError, GibbsError, GibbsRouletteError, GibbsShortSequenceError,
                   GibbsSymbolError, GibbsUnsupportedError)

__all__ = ["_abi", "GibbsError", "GibbsArgumentError", "GibbsCudaError", "GibbsRouletteError",
           "GibbsShortSequenceError", "GibbsSymbolError", "GibbsUnsupportedError"]
