"""otto_recommender_b200 -- B200-native co-visitation (co-event) counting engine.

A drop-in for ONE stage of nicolaivicol/otto-recommender: ``model/count_co_events.py`` (co-event
counting) and the per-aid top-N its consumer applies (``model/retrieve.py:41-51``).  Python host
code over a C-ABI library of hand-written sm_100a CUDA kernels (``libottocov.so``,
``include/ottocov.h``).  There is no CPU fallback: without the built library or without a CUDA
device every compute entry point raises.
"""
from .config import CoEventConfig, DEFAULT_CONFIG  # noqa: F401
from ._lib import OttocovError, load_library, library_path  # noqa: F401
from .engine import Engine, Table, WeightedTable  # noqa: F401

__all__ = ["CoEventConfig", "DEFAULT_CONFIG", "Engine", "Table", "WeightedTable", "OttocovError", "load_library",
           "library_path"]
