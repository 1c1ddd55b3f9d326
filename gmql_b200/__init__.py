"""gmql_b200 — B200 (sm_100a) back end for GMQL's region-operator hot path (MAP, COVER family, JOIN).

The product is `libgmqlcuda.so` (hand-written CUDA behind the C ABI of include/gmql_cuda.h).  This package is the
host-side mirror of the reference's operator interface (GMQLSparkExecutor.implement_rd dispatch for
IRGenometricMap / IRRegionCover / IRGenometricJoin), used by the parity tests and the benchmark; the production
host side is the Scala/JNI shim described in INTEGRATION.md.  There is no CPU fallback: importing works without
a GPU, calling an operator without one raises.
"""
from .regions import RegionDataset, STRAND_CODE, STRAND_CHAR  # noqa: F401
from .params import (Agg, CoverFlag, CoverParam, N, ANY, ALL, DistLess, DistGreater, MinDistance, Upstream,  # noqa: F401
                     DownStream, RegionBuilder, BinSize)
from .executor import GMQLCudaExecutor  # noqa: F401
from ._lib import GmqlCudaError, load_library  # noqa: F401

__all__ = ["RegionDataset", "GMQLCudaExecutor", "Agg", "CoverFlag", "CoverParam", "N", "ANY", "ALL", "DistLess",
           "DistGreater", "MinDistance", "Upstream", "DownStream", "RegionBuilder", "BinSize", "GmqlCudaError",
           "load_library"]
