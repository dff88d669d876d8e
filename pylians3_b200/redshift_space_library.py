"""Drop-in for Pylians3's `redshift_space_library`: `pos_redshift_space`.

Mirrors library/redshift_space_library/redshift_space_library.pyx:29-46: same name, argument order and in-place
semantics (`pos` is modified, nothing is returned).  It is the step the reference's snapshot drivers run before
MAS_library.MA (Pk_library/Pk_snapshot.py:60-64, MAS_library/MAS_gadget.py), so keeping it on the device lets
positions -> redshift space -> MA -> delta -> Pk run without host round trips.  The arithmetic runs in
libpyl_b200.so (pyl_pos_redshift_space); there is no CPU path.

NumPy float32 arrays are copied to the GPU and back; torch CUDA float32 tensors are used in place (zero-copy).
"""
import numpy as np
import torch

from . import _device as D
from . import _lib as L

__all__ = ["pos_redshift_space"]


def pos_redshift_space(pos, vel, BoxSize, Hubble, redshift, axis):
    """s = r + (1+z)/H(z) * v along `axis`, wrapped into the box; `pos` is updated in place."""
    D.require_cuda()
    if pos.ndim != 2 or pos.shape[1] != 3 or tuple(vel.shape) != tuple(pos.shape):
        raise ValueError("pos and vel must both be (N,3) float32 arrays")
    if axis not in (0, 1, 2):
        raise ValueError("axis must be 0, 1 or 2")
    dev = D.pick_device(pos, vel)
    inplace = D.is_cuda_tensor(pos) and pos.is_contiguous()
    if inplace and pos.dtype != torch.float32:
        raise ValueError("pos must be float32, got %s" % pos.dtype)
    pos_d = pos if inplace else D.to_device_f32(pos, dev, "pos")[0]
    vel_d, _ = D.to_device_f32(vel, dev, "vel")
    with torch.cuda.device(dev):
        L.check(L.load().pyl_pos_redshift_space(D.ptr(pos_d), D.ptr(vel_d), pos_d.shape[0],
                                                float(np.float32(BoxSize)), float(np.float32(Hubble)),
                                                float(np.float32(redshift)), int(axis), D.stream_ptr(dev)),
                "pyl_pos_redshift_space")
    if not inplace:
        if isinstance(pos, torch.Tensor):
            pos.copy_(pos_d)
        else:
            pos[...] = pos_d.cpu().numpy()
