"""Drop-in alias: ``import svox_t`` resolves to the B200-native implementation (svox_t_b200).

Lets code written against HaiminLuo/svox_t (``from svox_t import N3Tree, VolumeRenderer, Rays, ...``) run unchanged on
the sm_100a kernels. Only the hot path described in DESIGN.md is implemented; anything else raises."""
from svox_t_b200 import *            # noqa: F401,F403
from svox_t_b200 import __all__, __version__, csrc, helpers, p2v, renderer, svox  # noqa: F401
