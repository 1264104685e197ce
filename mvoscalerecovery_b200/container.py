"""Packed on-disk frame container and the reader of the reference's own hand-off file.

The reference hands the online front-end's output to the offline scale recovery as a pickled dict of ragged Python
lists, ``{'motions', 'feature3ds', 'feature2ds', 'move_flags'}``, written with ``np.save`` (src/main.py:149-154) and read
back with ``np.load(allow_pickle=True).item()`` (src/main_offline.py:27-32,52-55).  Unpickling thousands of small arrays
is far slower than the GPU path consumes them, so the batch path uses one flat little-endian file that maps 1:1 onto the
CSR structure-of-arrays frame batch of include/mvosr.h and can be memory-mapped and copied to the device without a
per-frame loop:

    bytes 0..63   header: magic b"MVOSRPK1", uint32 version (1), uint32 flags, uint64 n_frames F, uint64 n_features M, zero pad
    then, each section aligned to 64 bytes:
      offsets     int32  [F+1]
      move_flags  uint8  [F]
      motions     float64[F][12]      row-major [R|t] of the relative motion of every frame
      x, y, z     float32[M] each     triangulated features (camera frame, VO units)
      u, v        float32[M] each     their pixel coordinates

Host-side plumbing only (no arithmetic): nothing here needs the GPU.
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = b"MVOSRPK1"
VERSION = 1
_HEADER = struct.Struct("<8sIIQQ")
_ALIGN = 64
_SECTIONS = (("offsets", np.int32), ("move_flags", np.uint8), ("motions", np.float64),
             ("x", np.float32), ("y", np.float32), ("z", np.float32), ("u", np.float32), ("v", np.float32))


def _layout(F: int, M: int):
    shapes = dict(offsets=(F + 1,), move_flags=(F,), motions=(F, 12), x=(M,), y=(M,), z=(M,), u=(M,), v=(M,))
    pos, out = _ALIGN, {}
    for name, dt in _SECTIONS:
        n = int(np.prod(shapes[name])) * np.dtype(dt).itemsize
        out[name] = (pos, shapes[name], np.dtype(dt))
        pos = (pos + n + _ALIGN - 1) // _ALIGN * _ALIGN
    return out, pos


def pack_sequence(motions, feature3ds, feature2ds, move_flags):
    """Ragged per-frame lists (the reference's dict entries) -> dict of flat arrays in the container's layout."""
    F = len(move_flags)
    if not (len(motions) == len(feature3ds) == len(feature2ds) == F):
        raise ValueError("motions, feature3ds, feature2ds and move_flags must have one entry per frame")
    sizes = [int(np.asarray(a).reshape(-1, 3).shape[0]) for a in feature3ds]
    for f, (a, b) in enumerate(zip(sizes, feature2ds)):
        if np.asarray(b).reshape(-1, 2).shape[0] != a:
            raise ValueError("frame %d: feature3d and feature2d disagree in length" % f)
    off = np.zeros(F + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    f3 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 3) for a in feature3ds], 0) if F else np.zeros((0, 3), np.float32)
    f2 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 2) for a in feature2ds], 0) if F else np.zeros((0, 2), np.float32)
    mot = np.zeros((F, 12), np.float64)
    for f, m in enumerate(motions):
        mot[f] = np.asarray(m, np.float64).reshape(-1)[:12]
    c = np.ascontiguousarray
    return dict(offsets=off, move_flags=np.asarray(move_flags).astype(bool).astype(np.uint8), motions=mot,
                x=c(f3[:, 0]), y=c(f3[:, 1]), z=c(f3[:, 2]), u=c(f2[:, 0]), v=c(f2[:, 1]))


def save_packed(path: str, seq: dict) -> None:
    F, M = int(seq["move_flags"].shape[0]), int(seq["x"].shape[0])
    layout, total = _layout(F, M)
    with open(path, "wb") as fh:
        fh.write(_HEADER.pack(MAGIC, VERSION, 0, F, M).ljust(_ALIGN, b"\0"))
        for name, _ in _SECTIONS:
            pos, shape, dt = layout[name]
            a = np.ascontiguousarray(seq[name], dtype=dt)
            if a.shape != tuple(shape):
                raise ValueError("%s has shape %s, expected %s" % (name, a.shape, shape))
            fh.seek(pos)
            fh.write(a.astype(dt.newbyteorder("<"), copy=False).tobytes())
        fh.truncate(total)


def load_packed(path: str, mmap: bool = True) -> dict:
    """Dict of numpy arrays (read-only memory maps by default) in the container's layout."""
    with open(path, "rb") as fh:
        head = fh.read(_ALIGN)
    if len(head) < _HEADER.size:
        raise ValueError("%s: not an MVOSR container (too short)" % path)
    magic, version, _flags, F, M = _HEADER.unpack(head[:_HEADER.size])
    if magic != MAGIC:
        raise ValueError("%s: not an MVOSR container (bad magic)" % path)
    if version != VERSION:
        raise ValueError("%s: container version %d is not supported" % (path, version))
    layout, total = _layout(F, M)
    if os.path.getsize(path) < total:
        raise ValueError("%s: truncated container (%d bytes, expected %d)" % (path, os.path.getsize(path), total))
    out = {}
    for name, _ in _SECTIONS:
        pos, shape, dt = layout[name]
        if mmap:
            out[name] = np.memmap(path, dtype=dt.newbyteorder("<"), mode="r", offset=pos, shape=tuple(shape))
        else:
            with open(path, "rb") as fh:
                fh.seek(pos)
                out[name] = np.frombuffer(fh.read(int(np.prod(shape)) * dt.itemsize), dtype=dt.newbyteorder("<")).reshape(shape)
    if int(out["offsets"][0]) != 0 or int(out["offsets"][-1]) != M or np.any(np.diff(out["offsets"]) < 0):
        raise ValueError("%s: inconsistent offsets" % path)
    return out


def load_reference_npy(path: str) -> dict:
    """The reference's pickled hand-off (src/main.py:149-154), read the way src/main_offline.py:27-32 reads it."""
    data = np.load(path, allow_pickle=True).item()
    return pack_sequence(data["motions"], data["feature3ds"], data["feature2ds"], data["move_flags"])


def unpack_sequence(seq: dict) -> dict:
    """Back to the reference's dict of ragged lists (float64 arrays, as the reference's front-end produces them)."""
    off = np.asarray(seq["offsets"])
    f3 = np.stack([seq["x"], seq["y"], seq["z"]], 1).astype(np.float64)
    f2 = np.stack([seq["u"], seq["v"]], 1).astype(np.float64)
    F = off.shape[0] - 1
    return dict(motions=[np.array(seq["motions"][f]) for f in range(F)],
                feature3ds=[f3[off[f]:off[f + 1]] for f in range(F)], feature2ds=[f2[off[f]:off[f + 1]] for f in range(F)],
                move_flags=[bool(b) for b in seq["move_flags"]])


def result_prefix(data_path: str) -> str:
    """The prefix src/main_offline.py:37 derives from its input file name (fourth dot-separated field from the end, directory
    stripped, plus ``_``): ``result/00_result.npy.test_001.npy`` -> ``00_result_``."""
    return data_path.split('.')[-4].split('/')[-1] + '_'
