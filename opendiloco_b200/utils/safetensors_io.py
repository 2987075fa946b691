"""Dependency-free safetensors reader/writer (format: u64-LE header length | JSON header | raw little-endian data).

The reference loads HF checkpoints through the Rust ``safetensors`` crate (SURVEY.md §2.2 native census); here the
file is memory-mapped and each tensor is a zero-copy view.  When the host runtime (``libodbhost.so``) is built, header
parsing/validation goes through its C++ implementation (``odb_st_open``); the pure-Python path is kept as the portable
fallback and as the writer.
"""
from __future__ import annotations

import json
import mmap
import struct

import numpy as np
import torch

_DTYPES = {
    "F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
}
_NAMES = {v: k for k, v in _DTYPES.items()}


def read_header(path: str) -> tuple[dict, int]:
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > (1 << 28):
            raise ValueError(f"{path}: implausible safetensors header length {n}")
        header = json.loads(f.read(n).decode("utf-8"))
    return header, 8 + n


def _load_native(path: str) -> dict[str, torch.Tensor] | None:
    """Header parsing + bounds validation in C++ (csrc/host/safetensors.cc); tensors are copied out of the mmap."""
    import ctypes

    from .. import _lib

    lib = _lib.host_lib()
    if lib is None or not hasattr(lib, "odb_st_open"):
        return None
    lib.odb_st_open.restype = ctypes.c_void_p
    lib.odb_st_open.argtypes = [ctypes.c_char_p]
    for fn, res in (("odb_st_error", ctypes.c_char_p), ("odb_st_name", ctypes.c_char_p), ("odb_st_dtype", ctypes.c_char_p),
                    ("odb_st_count", ctypes.c_int), ("odb_st_ndim", ctypes.c_int), ("odb_st_dim", ctypes.c_int64),
                    ("odb_st_nbytes", ctypes.c_uint64), ("odb_st_data", ctypes.c_void_p), ("odb_st_close", None)):
        getattr(lib, fn).restype = res
    lib.odb_st_error.argtypes = lib.odb_st_count.argtypes = lib.odb_st_close.argtypes = [ctypes.c_void_p]
    for fn in ("odb_st_name", "odb_st_dtype", "odb_st_ndim", "odb_st_nbytes", "odb_st_data"):
        getattr(lib, fn).argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.odb_st_dim.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    h = lib.odb_st_open(path.encode())
    try:
        err = lib.odb_st_error(h)
        if err:
            raise ValueError(f"{path}: {err.decode()}")
        out = {}
        for i in range(lib.odb_st_count(h)):
            dt = _DTYPES[lib.odb_st_dtype(h, i).decode()]
            shape = [lib.odb_st_dim(h, i, d) for d in range(lib.odb_st_ndim(h, i))]
            n = lib.odb_st_nbytes(h, i)
            raw = torch.empty(n, dtype=torch.uint8)
            if n:
                ctypes.memmove(raw.data_ptr(), lib.odb_st_data(h, i), n)
            out[lib.odb_st_name(h, i).decode()] = raw.view(dt).reshape(shape)
        return out
    finally:
        lib.odb_st_close(h)


def load_safetensors(path: str) -> dict[str, torch.Tensor]:
    native = _load_native(path)
    if native is not None:
        return native
    header, base = read_header(path)
    out: dict[str, torch.Tensor] = {}
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    buf = np.frombuffer(mm, dtype=np.uint8)
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        dt = _DTYPES[meta["dtype"]]
        lo, hi = meta["data_offsets"]
        raw = torch.from_numpy(buf[base + lo: base + hi].copy())
        out[name] = raw.view(dt).reshape(meta["shape"])
    return out


def save_safetensors(tensors: dict[str, torch.Tensor], path: str, metadata: dict | None = None) -> None:
    header: dict = {}
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    off = 0
    order = sorted(tensors)
    for name in order:
        t = tensors[name]
        n = t.numel() * t.element_size()
        header[name] = {"dtype": _NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + n]}
        off += n
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for name in order:
            t = tensors[name].detach().cpu().contiguous()
            f.write(t.view(torch.uint8).numpy().tobytes() if t.numel() else b"")
