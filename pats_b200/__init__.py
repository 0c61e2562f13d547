"""pats_b200 -- B200-native (sm_100a) hot path of zju3dv/pats behind the reference's own call surface.

    pats_b200.modules        log_sinkhorn_iterations / log_optimal_transport / log_optimal_transport2
                             (reference models/modules.py:137-182)
    pats_b200.tensor_resize  tensor_resize(input, bound)          (reference setup/library.cpp:92-93)
    pats_b200.utils          origin_extract, Compute_imgs, ...    (reference utils/utils.py)
    pats_b200.install        install(): rebind those names inside the reference's modules so the
                             unmodified models/pats.py runs on the CUDA path
    pats_b200.host           host-buffer (numpy / CPU tensor) entry points of the C ABI
    pats_b200.dist           pair sharding + gather of match lists (torch.distributed)

All compute is in libpats_b200.so (hand-written CUDA, C ABI: include/pats_b200.h).  There is no
CPU fallback anywhere in this package.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["modules", "tensor_resize", "utils", "install", "host", "dist"]
