"""Minimal Kaldi ark/scp float-matrix I/O (host logic, numpy only).

The reference CLI writes code indices (`--indices_save_type ark`) and sub-quantizer embeddings (`--need_sub_quants true`)
with `kaldiio.WriteHelper("ark,scp,f:X.ark,X.scp")` (funcodec/bin/codec_inference.py:277-286) and reads embeddings for
`--run_mod decode_emb` through the `kaldi_ark` data type (funcodec/datasets/iterable_dataset.py `load_kaldi`).  kaldiio is
not a dependency here; this module writes / reads the same on-disk format: per entry `key SP \\0 B F M SP \\4 rows \\4 cols
float32-data` in the ark, and `key path:offset` in the scp (offset = position of the `\\0B` marker).
"""
import os
import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np


class ArkScpWriter:
    """`kaldiio.WriteHelper("ark,scp,f:<prefix>.ark,<prefix>.scp")` for 2-D float32 matrices."""

    def __init__(self, prefix: str):
        self.ark_path = prefix + ".ark"
        self._ark = open(self.ark_path, "wb")
        self._scp = open(prefix + ".scp", "wt", encoding="utf-8")

    def __call__(self, key: str, mat: np.ndarray) -> None:
        m = np.ascontiguousarray(np.asarray(mat, dtype="<f4"))
        if m.ndim != 2:
            raise ValueError("ArkScpWriter: matrix must be 2-D")
        self._ark.write(key.encode("utf-8") + b" ")
        offset = self._ark.tell()
        self._ark.write(b"\0BFM " + b"\4" + struct.pack("<i", m.shape[0]) + b"\4" + struct.pack("<i", m.shape[1]))
        self._ark.write(m.tobytes())
        self._ark.flush()                     # the "f" (flush) specifier of the reference's wspecifier
        self._scp.write(f"{key} {self.ark_path}:{offset}\n")
        self._scp.flush()

    def close(self) -> None:
        self._ark.close()
        self._scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_mat(spec: str) -> np.ndarray:
    """One `path:offset` scp value -> float32 (or float64) matrix."""
    path, _, off = spec.rpartition(":")
    with open(path, "rb") as f:
        f.seek(int(off))
        if f.read(2) != b"\0B":
            raise ValueError(f"{spec}: not a binary Kaldi matrix")
        tok = f.read(3)
        if tok not in (b"FM ", b"DM "):
            raise ValueError(f"{spec}: unsupported Kaldi type {tok!r} (float / double matrices only)")
        dt = "<f4" if tok == b"FM " else "<f8"
        assert f.read(1) == b"\4"
        rows = struct.unpack("<i", f.read(4))[0]
        assert f.read(1) == b"\4"
        cols = struct.unpack("<i", f.read(4))[0]
        data = np.frombuffer(f.read(rows * cols * np.dtype(dt).itemsize), dtype=dt)
    return data.reshape(rows, cols).astype(np.float32)


def read_scp_mats(scp_path: str) -> Iterator[Tuple[str, np.ndarray]]:
    with open(scp_path, "rt", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if line:
                key, spec = line.split(maxsplit=1)
                yield key, read_mat(spec)
