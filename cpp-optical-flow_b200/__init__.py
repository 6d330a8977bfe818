"""cpp-optical-flow_b200: the Horn-Schunck hot path of liuyang9609/Cpp-Optical-Flow as
hand-written sm_100a CUDA behind a C ABI (include/hs.h), plus the host-side mirror of the
reference's class.  The directory name has a hyphen; import it as `cpp_optical_flow_b200`
through the loader module of that name at the repository root."""
from . import _build, flowio, hs_ctypes, synth              # noqa: F401
from .hs_ctypes import HsError, load_library                # noqa: F401
from .horn_schunck import Solver, hornSchunck               # noqa: F401

build = _build.build
