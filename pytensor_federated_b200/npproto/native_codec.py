"""ctypes front-end of the native message codec (``csrc/codec.cu``).

``encode_arrays`` / ``decode_arrays`` handle a whole ``InputArrays`` / ``OutputArrays`` message in
one native call.  The pure-Python codec (:mod:`pytensor_federated_b200._pb`, ``rpc.py``) stays the
fallback and the oracle; :func:`available` tells whether the library can be used (it is built by
``pytensor_federated_b200.build`` and needs no GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

_MAX_DIMS = 16


class _PbItem(C.Structure):
    _fields_ = [
        ("data_off", C.c_longlong), ("data_len", C.c_longlong), ("dtype_off", C.c_longlong), ("dtype_len", C.c_longlong),
        ("ndim", C.c_int), ("n_strides", C.c_int), ("shape", C.c_longlong * _MAX_DIMS), ("strides", C.c_longlong * _MAX_DIMS),
    ]


_lib = None
_failed = False


def _load():
    global _lib, _failed
    if _lib is not None or _failed:
        return _lib
    try:
        from ..ops import native

        lib = native.load(build_if_missing=False)
        lib.b200_pb_encode_arrays.restype = C.c_longlong
        lib.b200_pb_encode_arrays.argtypes = [
            C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_char_p), C.POINTER(C.c_int),
            C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_char_p, C.c_void_p, C.c_longlong,
        ]
        lib.b200_pb_decode_arrays.restype = C.c_longlong
        lib.b200_pb_decode_arrays.argtypes = [
            C.c_void_p, C.c_longlong, C.POINTER(_PbItem), C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
        ]
        _lib = lib
    except Exception:  # noqa: BLE001 - library not built / not loadable: Python codec is used
        _failed = True
    return _lib


def available() -> bool:
    return _load() is not None


def encode_arrays(arrays: Sequence[np.ndarray], uuid: str = "") -> bytes:
    """Serialises arrays (C-contiguous copies are made when needed) as one Input/OutputArrays message."""
    lib = _load()
    if lib is None:
        raise RuntimeError("native codec unavailable")
    arrs = []
    for a in arrays:
        a = np.asarray(a)
        if a.dtype.hasobject:
            raise TypeError("object arrays are handled by the Python codec")
        if not a.flags.c_contiguous:
            a = a.copy(order="C")
        arrs.append(a)
    n = len(arrs)
    data = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
    nbytes = (C.c_longlong * n)(*[a.nbytes for a in arrs])
    dtypes = (C.c_char_p * n)(*[str(a.dtype).encode() for a in arrs])
    ndims = (C.c_int * n)(*[a.ndim for a in arrs])
    flat_shape = [d for a in arrs for d in a.shape]
    flat_strides = [s for a in arrs for s in a.strides]
    total_dims = max(1, len(flat_shape))
    shapes = (C.c_longlong * total_dims)(*flat_shape)
    strides = (C.c_longlong * total_dims)(*flat_strides)
    uuid_b = uuid.encode()
    size = lib.b200_pb_encode_arrays(n, data, nbytes, dtypes, ndims, shapes, strides, uuid_b, None, 0)
    buf = C.create_string_buffer(int(size))
    written = lib.b200_pb_encode_arrays(n, data, nbytes, dtypes, ndims, shapes, strides, uuid_b, buf, size)
    assert written == size
    return buf.raw


def decode_arrays(message: bytes) -> Tuple[List[np.ndarray], str]:
    """``(arrays, uuid)``; the arrays are read-only zero-copy views over ``message``."""
    lib = _load()
    if lib is None:
        raise RuntimeError("native codec unavailable")
    message = bytes(message)
    cap = 16
    uuid_off, uuid_len = C.c_longlong(), C.c_longlong()
    while True:
        items = (_PbItem * cap)()
        n = lib.b200_pb_decode_arrays(message, len(message), items, cap, C.byref(uuid_off), C.byref(uuid_len))
        if n == -2:
            raise TypeError("arrays with more than 16 dimensions are handled by the Python codec")
        if n < 0:
            raise ValueError("malformed ArraysToArrays message")
        if n <= cap:
            break
        cap = int(n)
    out = []
    for it in items[: int(n)]:
        dtype = np.dtype(message[it.dtype_off : it.dtype_off + it.dtype_len].decode())
        if dtype.hasobject:
            raise TypeError("object arrays are handled by the Python codec")
        shape = tuple(it.shape[: it.ndim])
        strides = tuple(it.strides[: it.n_strides]) if it.n_strides == it.ndim else None
        # the view is built over the item's OWN data field, so NumPy checks shape x strides against its length
        # (a short or inconsistent `data` raises instead of reading the neighbouring fields of the message)
        field = memoryview(message)[int(it.data_off) : int(it.data_off) + int(it.data_len)]
        try:
            out.append(np.ndarray(shape=shape, dtype=dtype, buffer=field, strides=strides))
        except (TypeError, ValueError) as ex:   # "buffer is too small" / "strides is incompatible with ... size of buffer"
            raise ValueError(f"malformed ndarray item: {ex}") from ex
    uuid = message[uuid_off.value : uuid_off.value + uuid_len.value].decode()
    return out, uuid
