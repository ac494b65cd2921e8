"""Compact binary container for RVQ code indices -- host logic, numpy only.

The reference ships `funcodec/modules/quantization/binary.py` (BitPacker / BitUnpacker) and an arithmetic coder, but its own
self-test fails in this checkout (`BitPacker.push` receives float values; `binary.test()` raises `ValueError: bytes must be in
range(0, 256)`), and nothing on the inference path calls it: `codecs.txt` is JSON, ~4.5 bytes per index.  This module defines
the wire format that replaces it: fixed-width little-endian bit packing of ceil(log2(K)) bits per index (10 bits for the
1024-entry codebooks = 1.25 bytes per index, 3.6x smaller than the JSONL), one record per utterance, stream-appendable.

Record layout (all little-endian):
    magic   4 bytes  b"FCB1"
    n_q     uint16   quantizer stages stored
    bits    uint8    bits per index (1..16)
    klen    uint8    length of the utterance key in bytes (UTF-8)
    frames  uint32   T'
    key     klen bytes
    payload ceil(n_q * frames * bits / 8) bytes: indices in [stage][frame] order, index i occupies bits
            [i * bits, (i + 1) * bits) of the payload viewed as one little-endian integer (LSB first, like the reference's
            BitPacker: `current_value += value << current_bits`, binary.py:70-80)
"""
import struct
from typing import BinaryIO, Iterator, Tuple

import numpy as np

MAGIC = b"FCB1"


def bits_for_codebook(codebook_size: int) -> int:
    return max(1, int(np.ceil(np.log2(codebook_size))))


def pack_indices(codes: np.ndarray, bits: int) -> bytes:
    """codes: [n_q, T'] non-negative ints < 2**bits -> payload bytes."""
    c = np.ascontiguousarray(codes, dtype=np.uint32).reshape(-1)
    if c.size and int(c.max()) >= (1 << bits):
        raise ValueError(f"index {int(c.max())} does not fit {bits} bits")
    shifts = np.arange(bits, dtype=np.uint32)
    bitmat = ((c[:, None] >> shifts[None, :]) & 1).astype(np.uint8)          # LSB first
    return np.packbits(bitmat.reshape(-1), bitorder="little").tobytes()


def unpack_indices(payload: bytes, n_q: int, frames: int, bits: int) -> np.ndarray:
    n = n_q * frames
    raw = np.unpackbits(np.frombuffer(payload, dtype=np.uint8), bitorder="little")[: n * bits]
    vals = (raw.reshape(n, bits).astype(np.int64) << np.arange(bits, dtype=np.int64)[None, :]).sum(axis=1)
    return vals.reshape(n_q, frames)


def write_record(f: BinaryIO, key: str, codes: np.ndarray, codebook_size: int = 1024) -> int:
    """Append one utterance ([n_q, T'] indices).  Returns the number of bytes written."""
    codes = np.asarray(codes)
    if codes.ndim != 2:
        raise ValueError("codes must be [n_q, T']")
    if codes.size and int(codes.min()) < 0:
        raise ValueError("negative index")
    bits = bits_for_codebook(codebook_size)
    kb = key.encode("utf-8")
    if len(kb) > 255:
        raise ValueError("key longer than 255 bytes")
    payload = pack_indices(codes, bits)
    head = MAGIC + struct.pack("<HBBI", codes.shape[0], bits, len(kb), codes.shape[1])
    f.write(head + kb + payload)
    return len(head) + len(kb) + len(payload)


def read_records(f: BinaryIO) -> Iterator[Tuple[str, np.ndarray]]:
    """Yields (key, codes [n_q, T'] int64) until end of file; raises on a truncated or foreign stream."""
    while True:
        head = f.read(12)
        if not head:
            return
        if len(head) != 12 or head[:4] != MAGIC:
            raise ValueError("not an FCB1 code stream (bad magic or truncated header)")
        n_q, bits, klen, frames = struct.unpack("<HBBI", head[4:])
        if not 1 <= bits <= 16:
            raise ValueError("corrupt header (bits)")
        key = f.read(klen)
        nbytes = (n_q * frames * bits + 7) // 8
        payload = f.read(nbytes)
        if len(key) != klen or len(payload) != nbytes:
            raise ValueError("truncated record")
        yield key.decode("utf-8"), unpack_indices(payload, n_q, frames, bits)


def codecs_txt_to_packed(txt_path: str, out_path: str, codebook_size: int = 1024) -> Tuple[int, int, int]:
    """Convert the reference's `codecs.txt` (JSONL) into the packed container.  Returns (utterances, text bytes, packed bytes)."""
    from .pipeline import parse_indices_line
    n = packed = text = 0
    with open(txt_path, "rt") as fin, open(out_path, "wb") as fout:
        for line in fin:
            if not line.strip():
                continue
            text += len(line.encode())
            key, arr = parse_indices_line(line)             # [T', n_q]
            packed += write_record(fout, key, arr.T, codebook_size)
            n += 1
    return n, text, packed


def packed_to_codecs_txt(packed_path: str, txt_path: str) -> int:
    import json
    n = 0
    with open(packed_path, "rb") as fin, open(txt_path, "wt") as fout:
        for key, codes in read_records(fin):
            fout.write(key + " " + json.dumps([codes.tolist()]) + "\n")
            n += 1
    return n
