"""seekdeep_b200 — B200-native hot path of SeekDeep's qluster (k-mer prefilter, affine-gap
global alignment with traceback, errorProfile) behind the C ABI in include/sdgpu.h.

Python here is harness glue (ctypes) mirroring the reference-side names; the compute lives in
csrc/*.cu (hand-written CUDA for sm_100a) and the host collapse loop in csrc/qluster_host.cpp.
"""
from . import _abi  # noqa: F401
from .scoring import GapScoringParameters, SubstituteMatrix, make_params  # noqa: F401
from .iterpars import CollapseIterations, IterPar  # noqa: F401
from .aligner import GpuAligner, SdgpuError  # noqa: F401
from .qluster import QlusterPars, population_cluster, qluster, qluster_many, shard_pair_blocks, shard_samples  # noqa: F401
