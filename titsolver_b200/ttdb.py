"""`.ttdb` particle storage: the reference's output format.

What `ParticleArray::write(time, series)` produces in the reference
(`tit/sph/particle_array.hpp:165-172` -> `tit/data/storage.cpp`): an SQLite
file with the tables `Settings / DataSeries / DataFrames / DataArrays`
(`storage.cpp:33-64`); one frame per output time, one array per varying field,
`type = (kind + 1) | rank << 8 | dim << 16` (`tit/data/type.hpp:176-189`),
`size` = number of elements, `data` = the packed little-endian elements as
Zstandard frames (`storage.cpp:351-372`, `core/serialization.hpp:57-128`).

This module reads and writes that format from Python (numpy arrays in, numpy
arrays out) so that results of the GPU path open in the reference's GUI /
exporters, and the reference's databases can be compared against. The C++
facade has the same thing behind the reference's names
(`include/tit_b200/data.hpp`). SQLite comes from the standard library,
Zstandard from `libzstd.so.1` through ctypes (no Python zstd module needed).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import sqlite3
import urllib.parse
from typing import Iterable

import numpy as np

# tit/data/type.hpp:30-42 — kind ids in storage order.
_KINDS = ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64")
RANK_SCALAR, RANK_VECTOR, RANK_MATRIX = 0, 1, 2

#: Field-set order of the WCSPH particle array (fluid_equations.hpp:41-48).
PARTICLE_FIELDS = ("m", "gamma", "grad_gamma", "rho", "drho_dt", "grad_rho", "p", "cs", "v", "dv_dt", "grad_v", "r", "dr", "L", "N", "phi", "rho_raw")

_SCHEMA = """
PRAGMA journal_mode = WAL;
PRAGMA foreign_keys = ON;
CREATE TABLE IF NOT EXISTS Settings (id INTEGER PRIMARY KEY CHECK (id = 0), max_series INTEGER) STRICT;
INSERT OR IGNORE INTO Settings (id, max_series) VALUES (0, 5);
CREATE TABLE IF NOT EXISTS DataSeries (id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT NOT NULL) STRICT;
CREATE TABLE IF NOT EXISTS DataFrames (
  id INTEGER PRIMARY KEY AUTOINCREMENT, series_id INTEGER NOT NULL, time REAL NOT NULL,
  FOREIGN KEY (series_id) REFERENCES DataSeries(id) ON DELETE CASCADE) STRICT;
CREATE TABLE IF NOT EXISTS DataArrays (
  id INTEGER PRIMARY KEY AUTOINCREMENT, frame_id INTEGER NOT NULL, name TEXT NOT NULL,
  type INTEGER, size INTEGER, data BLOB,
  FOREIGN KEY (frame_id) REFERENCES DataFrames(id) ON DELETE CASCADE) STRICT;
"""


def type_id(dtype, rank: int = RANK_SCALAR, dim: int = 1) -> int:
    """`Type::id()` (type.hpp:185-189); e.g. Vec<float64, 2> = 131338."""
    kind = _KINDS.index(np.dtype(dtype).name)
    if rank not in (RANK_SCALAR, RANK_VECTOR, RANK_MATRIX):
        raise ValueError(f"Invalid data type rank: {rank}.")
    if dim <= 0:
        raise ValueError(f"Dimensionality must be positive, but is {dim}.")
    if rank == RANK_SCALAR and dim != 1:
        raise ValueError(f"Dimensionality of a scalar must be 1, but is {dim}.")
    return (kind + 1) | rank << 8 | dim << 16


def decode_type(tid: int):
    """Inverse of `type_id` (type.hpp:176-180): `(dtype, rank, dim)`."""
    kind, rank, dim = (tid - 1) & 0xFF, (tid >> 8) & 0xFF, (tid >> 16) & 0xFF
    if kind >= len(_KINDS):
        raise ValueError(f"Invalid data kind ID: {kind}.")
    type_id(_KINDS[kind], rank, dim)  # validates rank / dim
    return np.dtype(_KINDS[kind]).newbyteorder("<"), rank, dim


def type_name(tid: int) -> str:
    """`Type::name()` (type.hpp:207-216)."""
    dt, rank, dim = decode_type(tid)
    kind = f"{dt.name}_t"
    return kind if rank == RANK_SCALAR else f"{'Vec' if rank == RANK_VECTOR else 'Mat'}<{kind}, {dim}>"


def _element_shape(rank: int, dim: int):
    return (dim,) * rank


class _Zstd:
    """The few libzstd entry points needed, bound on first use."""

    class _In(ctypes.Structure):
        _fields_ = [("src", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]

    class _Out(ctypes.Structure):
        _fields_ = [("dst", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            last = None
            for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
                if not name:
                    continue
                try:
                    lib = ctypes.CDLL(name)
                    break
                except OSError as e:  # pragma: no cover - depends on the host
                    last = e
            else:  # pragma: no cover
                raise ImportError(f"libzstd not found: {last}")
            lib.ZSTD_compressBound.restype = ctypes.c_size_t
            lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
            lib.ZSTD_compress.restype = ctypes.c_size_t
            lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            lib.ZSTD_isError.restype = ctypes.c_uint
            lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
            lib.ZSTD_getErrorName.restype = ctypes.c_char_p
            lib.ZSTD_getErrorName.argtypes = [ctypes.c_size_t]
            lib.ZSTD_createDStream.restype = ctypes.c_void_p
            lib.ZSTD_freeDStream.argtypes = [ctypes.c_void_p]
            lib.ZSTD_decompressStream.restype = ctypes.c_size_t
            lib.ZSTD_decompressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(cls._Out), ctypes.POINTER(cls._In)]
            cls._lib = lib
        return cls._lib

    @classmethod
    def compress(cls, raw: bytes, level: int = 3) -> bytes:
        lib = cls.lib()
        out = ctypes.create_string_buffer(lib.ZSTD_compressBound(len(raw)))
        n = lib.ZSTD_compress(out, len(out), raw, len(raw), level)
        if lib.ZSTD_isError(n):
            raise RuntimeError(f"ZSTD compression failed: {lib.ZSTD_getErrorName(n).decode()}.")
        return out.raw[:n]

    @classmethod
    def decompress(cls, packed: bytes, size: int) -> bytes:
        """All concatenated frames of `packed` (the reference streams them without a
        content size, core/zstd.cpp:53-100); the result must be `size` bytes."""
        lib = cls.lib()
        ds = lib.ZSTD_createDStream()
        try:
            src = ctypes.create_string_buffer(packed, len(packed))
            dst = ctypes.create_string_buffer(max(size, 1))
            inb = cls._In(ctypes.cast(src, ctypes.c_void_p), len(packed), 0)
            outb = cls._Out(ctypes.cast(dst, ctypes.c_void_p), size, 0)
            status = 0
            while inb.pos < inb.size:
                before = (inb.pos, outb.pos)
                status = lib.ZSTD_decompressStream(ds, ctypes.byref(outb), ctypes.byref(inb))
                if lib.ZSTD_isError(status):
                    raise RuntimeError(f"ZSTD decompression failed: {lib.ZSTD_getErrorName(status).decode()}.")
                if (inb.pos, outb.pos) == before:
                    break
            if inb.pos != inb.size or outb.pos != size:
                raise RuntimeError("ZSTD decompression failed: data size mismatch.")
            if status != 0:
                raise RuntimeError("ZSTD decompression failed: truncated frame.")
            return dst.raw[:size]
        finally:
            lib.ZSTD_freeDStream(ds)


class Array:
    """`ArrayView` (storage.hpp:46-144)."""

    def __init__(self, storage: "Storage", array_id: int):
        self.storage, self.id = storage, array_id

    def __eq__(self, other):
        return isinstance(other, Array) and other.id == self.id

    def __hash__(self):
        return hash(("array", self.id))

    def _col(self, col):
        row = self.storage._db.execute(f"SELECT {col} FROM DataArrays WHERE id = ?", (self.id,)).fetchone()
        if row is None:
            raise KeyError("Invalid array ID!")
        return row[0]

    @property
    def name(self) -> str:
        return self._col("name")

    @property
    def type(self) -> int:
        return self._col("type")

    @property
    def size(self) -> int:
        return self._col("size")

    def write(self, values, rank: int | None = None) -> None:
        """Replace the contents. `values`: (n,) scalars, (n, D) vectors or (n, D, D)
        matrices (rows packed, `serialization.hpp:118-128`); an (n, D) or (n, D, D)
        array may be declared as something else with `rank`."""
        a = np.ascontiguousarray(values)
        if rank is None:
            rank = a.ndim - 1
        dim = 1 if rank == RANK_SCALAR else a.shape[-1]
        if a.shape[1:] != _element_shape(rank, dim):
            raise ValueError("Data size mismatch!")
        tid = type_id(a.dtype, rank, dim)
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        with self.storage._db:
            cur = self.storage._db.execute("UPDATE DataArrays SET type = ?, size = ?, data = ? WHERE id = ?", (tid, a.shape[0], _Zstd.compress(raw), self.id))
        if cur.rowcount != 1:
            raise KeyError("Invalid array ID!")

    def read(self) -> np.ndarray:
        tid, size, blob = self.storage._db.execute("SELECT type, size, data FROM DataArrays WHERE id = ?", (self.id,)).fetchone()
        dt, rank, dim = decode_type(tid)
        shape = (size,) + _element_shape(rank, dim)
        raw = _Zstd.decompress(bytes(blob), int(np.prod(shape)) * dt.itemsize)
        return np.frombuffer(raw, dtype=dt).reshape(shape).copy()


class Frame:
    """`FrameView` (storage.hpp:148-226)."""

    def __init__(self, storage: "Storage", frame_id: int):
        self.storage, self.id = storage, frame_id

    def __eq__(self, other):
        return isinstance(other, Frame) and other.id == self.id

    def __hash__(self):
        return hash(("frame", self.id))

    @property
    def time(self) -> float:
        row = self.storage._db.execute("SELECT time FROM DataFrames WHERE id = ?", (self.id,)).fetchone()
        if row is None:
            raise KeyError("Invalid frame ID!")
        return row[0]

    @property
    def num_arrays(self) -> int:
        return self.storage._db.execute("SELECT COUNT(*) FROM DataArrays WHERE frame_id = ?", (self.id,)).fetchone()[0]

    def arrays(self) -> list[Array]:
        return [Array(self.storage, i) for (i,) in self.storage._db.execute("SELECT id FROM DataArrays WHERE frame_id = ? ORDER BY id ASC", (self.id,))]

    def find_array(self, name: str) -> Array | None:
        row = self.storage._db.execute("SELECT id FROM DataArrays WHERE frame_id = ? AND name = ?", (self.id, name)).fetchone()
        return Array(self.storage, row[0]) if row else None

    def create_array(self, name: str) -> Array:
        if not name:
            raise ValueError("Array name must not be empty!")
        if not self.storage.check_frame(self):
            raise KeyError("Invalid frame ID!")
        if self.find_array(name) is not None:
            raise ValueError("Array already exists!")
        with self.storage._db:
            cur = self.storage._db.execute("INSERT INTO DataArrays (frame_id, name) VALUES (?, ?)", (self.id, name))
        return Array(self.storage, cur.lastrowid)

    def read(self) -> dict[str, np.ndarray]:
        """All arrays of the frame by name."""
        return {a.name: a.read() for a in self.arrays()}


class Series:
    """`SeriesView` (storage.hpp:229-310)."""

    def __init__(self, storage: "Storage", series_id: int):
        self.storage, self.id = storage, series_id

    def __eq__(self, other):
        return isinstance(other, Series) and other.id == self.id

    def __hash__(self):
        return hash(("series", self.id))

    @property
    def name(self) -> str:
        row = self.storage._db.execute("SELECT name FROM DataSeries WHERE id = ?", (self.id,)).fetchone()
        if row is None:
            raise KeyError("Invalid series ID!")
        return row[0] or ""

    @property
    def num_frames(self) -> int:
        return self.storage._db.execute("SELECT COUNT(*) FROM DataFrames WHERE series_id = ?", (self.id,)).fetchone()[0]

    def frames(self) -> list[Frame]:
        return [Frame(self.storage, i) for (i,) in self.storage._db.execute("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id ASC", (self.id,))]

    def frame(self, index: int) -> Frame:
        row = self.storage._db.execute("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id ASC LIMIT 1 OFFSET ?", (self.id, index)).fetchone()
        if row is None:
            raise IndexError(f"Frame index '{index}' out of bounds.")
        return Frame(self.storage, row[0])

    def last_frame(self) -> Frame:
        row = self.storage._db.execute("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id DESC LIMIT 1", (self.id,)).fetchone()
        if row is None:
            raise IndexError("Series is empty!")
        return Frame(self.storage, row[0])

    def create_frame(self, time: float) -> Frame:
        """Frame times of a series increase (storage.cpp:224-226)."""
        if not self.storage.check_series(self):
            raise KeyError("Invalid series ID!")
        if self.num_frames and not time > self.last_frame().time:
            raise ValueError("Frame time must be greater than the last frame time!")
        with self.storage._db:
            cur = self.storage._db.execute("INSERT INTO DataFrames (series_id, time) VALUES (?, ?)", (self.id, float(time)))
        return Frame(self.storage, cur.lastrowid)

    def write_particles(self, time: float, fields: dict[str, np.ndarray], names: Iterable[str] | None = None) -> Frame:
        """`ParticleArray::write` (particle_array.hpp:165-172): one frame, one array per
        field in the field-set order. Matrix fields `(n, D, D)` are told from vector
        fields by their shape."""
        frame = self.create_frame(time)
        for name in (names if names is not None else [f for f in PARTICLE_FIELDS if f in fields]):
            frame.create_array(name).write(fields[name])
        return frame


def _id(x) -> int:
    return x.id if hasattr(x, "id") else int(x)


class Storage:
    """`data::Storage` (storage.hpp:314-520): `Storage(path)` opens or creates."""

    def __init__(self, path: str = ":memory:", read_only: bool = False):
        self.read_only = read_only
        if read_only:
            self._db = sqlite3.connect(f"file:{urllib.parse.quote(os.path.abspath(str(path)))}?mode=ro", uri=True)
            self._db.execute("SELECT COUNT(*) FROM sqlite_master").fetchone()
        else:
            self._db = sqlite3.connect(str(path))
            self._db.executescript(_SCHEMA)
        self._db.execute("PRAGMA foreign_keys = ON")
        self.path = "" if str(path) == ":memory:" else str(path)

    def close(self) -> None:
        self._db.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # series
    @property
    def max_series(self) -> int:
        return self._db.execute("SELECT max_series FROM Settings").fetchone()[0]

    def set_max_series(self, value: int) -> None:
        """Lowering the cap drops the oldest series (storage.cpp:77-92)."""
        if value <= 0:
            raise ValueError("Maximum number of series must be positive!")
        with self._db:
            self._db.execute("UPDATE Settings SET max_series = ?", (value,))
            extra = self.num_series - value
            if extra > 0:
                self._db.execute("DELETE FROM DataSeries WHERE id IN (SELECT id FROM DataSeries ORDER BY id ASC LIMIT ?)", (extra,))

    @property
    def num_series(self) -> int:
        return self._db.execute("SELECT COUNT(*) FROM DataSeries").fetchone()[0]

    def series(self, index: int | None = None):
        if index is None:
            return [Series(self, i) for (i,) in self._db.execute("SELECT id FROM DataSeries ORDER BY id ASC")]
        row = self._db.execute("SELECT id FROM DataSeries ORDER BY id ASC LIMIT 1 OFFSET ?", (index,)).fetchone()
        if row is None:
            raise IndexError(f"Series index '{index}' out of bounds.")
        return Series(self, row[0])

    def last_series(self) -> Series:
        row = self._db.execute("SELECT id FROM DataSeries ORDER BY id DESC LIMIT 1").fetchone()
        if row is None:
            raise IndexError("No series in the storage!")
        return Series(self, row[0])

    def create_series(self, name: str = "") -> Series:
        """At the cap the oldest series is evicted first (storage.cpp:130-143)."""
        with self._db:
            if self.num_series >= self.max_series:
                self._db.execute("DELETE FROM DataSeries WHERE id IN (SELECT id FROM DataSeries ORDER BY id ASC LIMIT 1)")
            cur = self._db.execute("INSERT INTO DataSeries (name) VALUES (?)", (name,))
        return Series(self, cur.lastrowid)

    def delete_series(self, series) -> None:
        with self._db:
            self._db.execute("DELETE FROM DataSeries WHERE id = ?", (_id(series),))

    def check_series(self, series) -> bool:
        return self._db.execute("SELECT id FROM DataSeries WHERE id = ?", (_id(series),)).fetchone() is not None

    # frames, arrays
    def delete_frame(self, frame) -> None:
        with self._db:
            self._db.execute("DELETE FROM DataFrames WHERE id = ?", (_id(frame),))

    def check_frame(self, frame) -> bool:
        return self._db.execute("SELECT id FROM DataFrames WHERE id = ?", (_id(frame),)).fetchone() is not None

    def delete_array(self, array) -> None:
        with self._db:
            self._db.execute("DELETE FROM DataArrays WHERE id = ?", (_id(array),))

    def check_array(self, array) -> bool:
        return self._db.execute("SELECT id FROM DataArrays WHERE id = ?", (_id(array),)).fetchone() is not None


def write_solver_frame(series: Series, time: float, solver, names: Iterable[str] = PARTICLE_FIELDS) -> Frame:
    """Download the fields of a `titsolver_b200.Solver` (original particle order) and
    store them as one frame — the Python counterpart of `particles.write(time, series)`
    in the reference's time loop (wcsph.cpp:160, 185-188)."""
    return series.write_particles(time, {f: solver.download(f) for f in names}, names=list(names))
