"""fempy_b200 -- B200-native (sm_100a) assembly hot path behind FEMpy's Model/Problem API.

The product is the CUDA library `csrc/libfemb200.so` (C ABI: include/femb200.h); this package is
the thin Python host side: the ctypes binding, the assembly mixin for the reference's FEMpyProblem, and
stand-alone table generators / element and constitutive descriptors for use without the reference.  There is no CPU fallback.
"""
from . import assembly_utils as AssemblyUtils  # noqa: F401
from . import constitutive as Constitutive  # noqa: F401
from . import elements as Elements  # noqa: F401
from . import meshes, tables  # noqa: F401
from .plan import AssemblyPlan  # noqa: F401
from .problem import B200AssemblyMixin, b200ProblemClass, install, materialFor, planFor  # noqa: F401

__version__ = "0.1.0"
