"""B200-native mip-pyramid generator: drop-in for the nvpro_pyramid hot path.

Only the path lives here: the C ABI binding (``_lib``), the host-side mirror of
``nvproCmdPyramidDispatch`` (``pyramid``) and the sharding helper for batches
(``batch``).  Importing fails loudly when libnvpyr.so has not been built.
"""
from .pyramid import *  # noqa: F401,F403
from . import pyramid, batch  # noqa: F401
