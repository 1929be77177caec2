"""ruviii_b200: B200-native (sm_100a) engine for the RUV-III hot path of RMolania/RUVIIIPRPS.

The product is ``lib/libruviii_b200.so`` (C ABI in ``include/ruviii.h``); this package is the thin host side:
``_cabi`` (ctypes), ``api`` (the reference's R interface restated in Python), ``synth`` (synthetic inputs)
and ``sharding`` (gene shards for one-process-per-GPU launches).  Importing the package never loads the
library; the first compute call does, and raises if it is missing.
"""
from .api import (ReplicateData, SummarizedExperiment, fastruvIII, get_engine, group_codes, normalise, ruvIII,  # noqa: F401
                  ruvIIIMultipleK, set_engine)
from .prps import (prpsForCategoricalUV, prpsForContinuousUV, supervisedPRPS,  # noqa: F401
                   replicate_data_from_metadata)
from .pca import computePCA  # noqa: F401
from .ncg import supervisedFindNCG  # noqa: F401
from ._cabi import Engine, RuviiiError, make_params, MODE_FAST, MODE_STACKED  # noqa: F401

__version__ = "0.1.0"
