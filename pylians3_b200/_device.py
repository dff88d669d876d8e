"""Device-buffer plumbing: torch tensors carry HBM allocations and the CUDA stream.

torch is used ONLY for memory (caching allocator, pinned staging) and stream handles; every
arithmetic step of the path runs in libpyl_b200.so.
"""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pylians3_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def is_cuda_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """Grow-only scratch buffer per (device, tag), owned by torch's caching allocator."""
    nbytes = int(nbytes)
    if nbytes <= 0:
        return None
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _ws_cache[key]
            del buf
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspaces():
    _ws_cache.clear()
    _pinned.clear()
    del _result_pool[:]


_pinned = {}


def to_host_numpy(t):
    """Device tensor -> NumPy array through a cached PINNED staging buffer (a pageable `.cpu()` of the few-MB
    accumulator buffer was measured at 0.8-2.9 ms; pinned it is PCIe-bound).  The returned array aliases the
    staging buffer: it is valid until the next call for the same device -- copy what must outlive that."""
    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    key = t.device.index
    buf = _pinned.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _pinned[key] = buf
    view = buf[:nbytes].view(t.dtype).view(t.shape)
    view.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return view.numpy()


_result_pool = []          # [pinned tensor, weakref to the ndarray handed out over it]


def result_block(nwords):
    """(pinned float64 tensor, float64 ndarray over the same memory) for one set of results.

    Blocks are recycled: a block is free again once the ndarray handed out with it -- the `.base` of every view
    the caller received -- has been garbage collected.  A steady-state loop therefore touches the same few
    blocks (no page faults, no cudaHostAlloc per call), while every result the caller keeps stays valid."""
    import weakref
    nwords = int(nwords)
    for entry in _result_pool:
        t, ref = entry
        if t.numel() >= nwords and (ref is None or ref() is None):
            root = t.numpy()
            entry[1] = weakref.ref(root)
            return t, root
    t = torch.empty(max(nwords, 1 << 16), dtype=torch.float64, pin_memory=True)
    root = t.numpy()
    _result_pool.append([t, weakref.ref(root)])
    return t, root


def to_device_f32(x, device, name):
    """numpy float32 array or torch tensor -> contiguous float32 CUDA tensor (zero-copy if possible).

    Returns (tensor, was_host).  dtype is checked, never converted silently: the reference's
    typed memoryviews reject anything but float32 as well (MAS_library.pyx:123)."""
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float32:
            raise ValueError("%s must be float32, got %s" % (name, x.dtype))
        if x.is_cuda:
            return (x if x.is_contiguous() else x.contiguous()), False
        return x.contiguous().to(device, non_blocking=True), True
    a = np.asarray(x)
    if a.dtype != np.float32:
        raise ValueError("%s must be float32 (Buffer dtype mismatch, expected 'float32_t' but got '%s')"
                         % (name, a.dtype))
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a).to(device, non_blocking=False), True


def pick_device(*objs):
    for o in objs:
        if is_cuda_tensor(o):
            return o.device
    return torch.device("cuda", torch.cuda.current_device())
