"""Tiny HDF5 writer for tests (no h5py in this image): ONE N-dimensional integer dataset in the root group, laid out
like libhdf5 / h5py write it -- superblock version 0, symbol-table root group (B-tree v1 + symbol node + local
heap), version-1 object headers, layout version 3 -- contiguous like h5py's `create_dataset(name, data=a)`
(tools/CNN_training/inference.py:454-455), or chunked through a v1 chunk B-tree with optional shuffle + deflate
like `create_dataset(..., chunks=..., compression="gzip", shuffle=True)`.  Follows the HDF5 File Format
Specification 3.0; apps/h5_reader.h is the consumer (tests/test_h5_reader.py; a libhdf5-written file is the
independent fixture there)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages: list[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def _datatype(dtype: np.dtype) -> bytes:
    dtype = np.dtype(dtype)
    assert dtype.kind in "iu"
    bits0 = (1 if dtype.byteorder == ">" else 0) | (8 if dtype.kind == "i" else 0)
    return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, 8 * dtype.itemsize)


def _dataspace(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def write_h5(path: str, name: str, array: np.ndarray, chunks=None, deflate: bool = False, shuffle: bool = False,
             userblock: int = 0):
    a = np.ascontiguousarray(array)
    out = bytearray()
    # --- fixed positions (relative to the base address = start of the superblock) ---
    SB = 96                       # superblock v0 with 8-byte offsets: 24 + 4*8 + 40
    heap_addr = SB                # local heap header (32 bytes) + data segment
    names = b"\0" * 8 + name.encode() + b"\0"
    heap_data = _pad8(names)
    heap_data += b"\0" * (88 - len(heap_data)) if len(heap_data) < 88 else b""
    heap_data_addr = heap_addr + 32
    btree_addr = heap_data_addr + len(heap_data)
    K_LEAF, K_INT = 4, 16
    btree_size = 8 + 16 + (2 * K_INT + 1) * 8 + 2 * K_INT * 8
    snod_addr = btree_addr + btree_size
    snod_size = 8 + 2 * K_LEAF * 40
    root_hdr_addr = snod_addr + snod_size
    root_hdr = _object_header([_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))])
    ds_hdr_addr = root_hdr_addr + len(root_hdr)

    # --- dataset header (its size does not depend on the addresses it holds) ---
    def dataset_header(data_addr, data_size, chunk_btree=UNDEF):
        msgs = [_message(0x01, _dataspace(a.shape)), _message(0x03, _datatype(a.dtype), flags=1)]
        if chunks is None:
            msgs.append(_message(0x08, struct.pack("<BBQQ", 3, 1, data_addr, data_size)))
        else:
            lay = struct.pack("<BBBQ", 3, 2, len(chunks) + 1, chunk_btree)
            lay += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", a.dtype.itemsize)
            msgs.append(_message(0x08, lay))
            filt = []
            if shuffle:
                filt.append(struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", a.dtype.itemsize) + b"\0" * 4)
            if deflate:
                filt.append(struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 6) + b"\0" * 4)
            if filt:
                msgs.append(_message(0x0B, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
        return _object_header(msgs)

    ds_hdr_len = len(dataset_header(0, 0, 0))
    data_addr = ds_hdr_addr + ds_hdr_len
    if chunks is None:
        raw = a.tobytes()
        ds_hdr = dataset_header(data_addr, len(raw))
        tail = raw
    else:
        # one leaf node holds all chunks (fine for the handful a test needs)
        import itertools
        grid = [range(0, s, c) for s, c in zip(a.shape, chunks)]
        entries, blobs = [], []
        rank = len(chunks)
        node_size = 8 + 16 + (2 * 32 + 1) * (8 + 8 * (rank + 1)) + 2 * 32 * 8
        pos = data_addr + node_size
        for origin in itertools.product(*grid):
            block = np.zeros(chunks, dtype=a.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(origin, chunks, a.shape))
            block[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
            blob = block.tobytes()
            if shuffle:
                blob = np.frombuffer(blob, np.uint8).reshape(-1, a.dtype.itemsize).T.tobytes()
            if deflate:
                blob = zlib.compress(blob, 6)
            entries.append((len(blob), origin, pos))
            blobs.append(blob)
            pos += len(blob)
        assert len(entries) <= 64
        node = b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + struct.pack("<QQ", UNDEF, UNDEF)
        for size, origin, addr in entries:
            node += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in origin) + struct.pack("<Q", 0)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in a.shape) + struct.pack("<Q", 0)  # final key
        node += b"\0" * (node_size - len(node))
        ds_hdr = dataset_header(0, 0, data_addr)
        tail = node + b"".join(blobs)
    assert len(ds_hdr) == ds_hdr_len
    eof = data_addr + len(tail)

    # --- superblock ---
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, K_LEAF, K_INT, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, eof + userblock, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == SB
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF if True else 0, heap_data_addr)
    btree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
    btree += struct.pack("<Q", 0) + struct.pack("<Q", snod_addr) + struct.pack("<Q", 8)
    btree += b"\0" * (btree_size - len(btree))
    snod = b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", 8, ds_hdr_addr, 0, 0)
    snod += b"\0" * (snod_size - len(snod))
    out += b"\0" * userblock + sb + heap + heap_data + btree + snod + root_hdr + ds_hdr + tail
    with open(path, "wb") as f:
        f.write(bytes(out))


if __name__ == "__main__":
    import sys
    write_h5(sys.argv[1], "nlogprobs", np.load(sys.argv[2]))
