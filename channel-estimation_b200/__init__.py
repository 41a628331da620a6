"""B200-native Monte-Carlo hot path of rnissel/Channel-Estimation (host side).

The package directory name contains a hyphen; import it with
``importlib.import_module("channel-estimation_b200")`` or through the root-level alias module
``chest_b200``.  Everything numerical on the hot path runs in ``libchest_b200.so`` (CUDA, sm_100a);
there is no CPU fallback."""
from . import _lib                      # noqa: F401
from .context import DeviceContext, ChestError   # noqa: F401
