"""tamc -- Python (ctypes) binding of libtamc.so, the B200-native photon Monte-Carlo transport that
stands in for /root/reference/src/mcpolar.f90:151-173.

This is host-side plumbing only: every compute call goes through the C ABI declared in
include/tamc.h into hand-written CUDA (csrc/).  There is no CPU path here -- if libtamc.so has not
been built, importing succeeds but the first use raises, loudly.
"""
from .binding import (  # noqa: F401
    MCTransport, Stats, TamcError, RECORD_DTYPE, SCATTER, FRESNEL, PERIODIC, device_count, lib, lib_path, comm_unique_id,
    pin_host, unpin_host,
)
from .mcgrid import gridset, init_opt1, delta_for  # noqa: F401
from . import configs  # noqa: F401
