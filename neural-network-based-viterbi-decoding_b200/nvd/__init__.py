"""Host layer over libnvd.so (ctypes).  No CPU compute lives here: every operation goes to the
CUDA library and raises NvdError if it is missing or fails."""
from ._lib import NvdError, load  # noqa: F401
