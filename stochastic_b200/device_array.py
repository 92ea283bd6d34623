"""Result arrays: thin numpy-like view of a torch CUDA tensor.

The reference returns JAX device arrays which callers index, slice, reshape, feed to numpy /
`jnp` functions and call `.block_until_ready()` on (reference
`benchmarks/benchmark_sde_solver.py:29`, `tests/test_kloeden_platen_statistics.py:29-31`).
`DeviceArray` offers the same surface over a tensor that stays in HBM until it is actually
needed on the host; `__dlpack__` hands it to any other framework without a copy.
"""
import numpy as np


class DeviceArray:
    __array_priority__ = 50

    def __init__(self, tensor):
        self._t = tensor

    # -- interop ---------------------------------------------------------------------------
    def torch(self):
        return self._t

    def numpy(self):
        return self._t.detach().cpu().numpy()

    __sdeb200_host__ = numpy

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __dlpack__(self, *a, **k):
        return self._t.__dlpack__(*a, **k)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()

    def block_until_ready(self):
        import torch
        if self._t.is_cuda:
            torch.cuda.current_stream(self._t.device).synchronize()
        return self

    # -- numpy-like surface ------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.ndim

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    @property
    def T(self):
        return DeviceArray(self._t.permute(*reversed(range(self._t.ndim))))

    def __len__(self):
        return self._t.shape[0]

    def __getitem__(self, idx):
        idx = _unwrap_index(idx)
        return DeviceArray(self._t[idx])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self._t.reshape(*shape))

    def transpose(self, *axes):
        if not axes:
            return self.T
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return DeviceArray(self._t.permute(*axes))

    def squeeze(self, axis=None):
        return DeviceArray(self._t.squeeze() if axis is None else self._t.squeeze(axis))

    def astype(self, dtype):
        return self.numpy().astype(dtype)

    def item(self):
        return self._t.item()

    def tolist(self):
        return self._t.tolist()

    def __float__(self):
        return float(self._t.item())

    def __int__(self):
        return int(self._t.item())

    def __bool__(self):
        return bool(self._t.item())

    def __repr__(self):
        return f"DeviceArray({self.numpy()!r}, device={self._t.device})"

    def _reduce(self, name, axis=None, **kw):
        t = self._t
        if name == "mean":
            r = t.mean() if axis is None else t.mean(dim=axis)
        elif name == "sum":
            r = t.sum() if axis is None else t.sum(dim=axis)
        elif name == "std":
            r = t.std(unbiased=False) if axis is None else t.std(dim=axis, unbiased=False)
        elif name == "max":
            r = t.max() if axis is None else t.max(dim=axis).values
        elif name == "min":
            r = t.min() if axis is None else t.min(dim=axis).values
        return DeviceArray(r)

    def mean(self, axis=None, **kw): return self._reduce("mean", axis)
    def sum(self, axis=None, **kw): return self._reduce("sum", axis)
    def std(self, axis=None, **kw): return self._reduce("std", axis)
    def max(self, axis=None, **kw): return self._reduce("max", axis)
    def min(self, axis=None, **kw): return self._reduce("min", axis)

    def _bin(self, other, fn, reverse=False):
        o = other._t if isinstance(other, DeviceArray) else other
        if isinstance(o, np.ndarray):
            import torch
            o = torch.as_tensor(o, device=self._t.device)
        return DeviceArray(fn(o, self._t) if reverse else fn(self._t, o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return DeviceArray(-self._t)
    def __abs__(self): return DeviceArray(self._t.abs())
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)


def _unwrap_index(idx):
    if isinstance(idx, DeviceArray):
        return idx._t
    if isinstance(idx, tuple):
        return tuple(_unwrap_index(i) for i in idx)
    return idx
