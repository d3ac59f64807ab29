"""Readers/writers for the reference's SRS files — rmp-serde msgpack of
  srs/pallas.srs, srs/vesta.srs            [g, h]                     points: bin8(33) compressed   (ipa.rs:53-75, serialization.rs:65-84)
  srs/test_pallas.srs, srs/test_vesta.srs  [g, h, {n: [[point]; n]}]  points: bin8(65) uncompressed (precomputed_srs.rs:35-51, serialization.rs:108-146)
The files are arrays of equally sized `bin8` records, so the point payloads are sliced out with one strided numpy view (no
per-element Python objects); decoding to affine Montgomery points happens on the device (Context.decompress_points /
Context.points_from_uncompressed).  Only the msgpack subset those files use is understood; anything else raises ValueError."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SrsFile:
    g: np.ndarray                      # uint8 [n, 33] (compressed) or [n, 65] (uncompressed)
    h: np.ndarray                      # uint8 [33] or [65]
    lagrange_bases: dict = field(default_factory=dict)   # n -> uint8 [n, chunks, 65]

    @property
    def compressed(self) -> bool:
        return self.g.shape[1] == 33


class _Reader:
    def __init__(self, buf: bytes):
        self.b = memoryview(buf)
        self.a = np.frombuffer(buf, dtype=np.uint8)
        self.p = 0

    def byte(self) -> int:
        v = self.b[self.p]
        self.p += 1
        return v

    def _be(self, fmt, size):
        v = struct.unpack_from(fmt, self.b, self.p)[0]
        self.p += size
        return v

    def array_len(self) -> int:
        t = self.byte()
        if 0x90 <= t <= 0x9F:
            return t & 0x0F
        if t == 0xDC:
            return self._be(">H", 2)
        if t == 0xDD:
            return self._be(">I", 4)
        raise ValueError(f"msgpack: expected an array at byte {self.p - 1}, found 0x{t:02x}")

    def map_len(self) -> int:
        t = self.byte()
        if 0x80 <= t <= 0x8F:
            return t & 0x0F
        if t == 0xDE:
            return self._be(">H", 2)
        if t == 0xDF:
            return self._be(">I", 4)
        raise ValueError(f"msgpack: expected a map at byte {self.p - 1}, found 0x{t:02x}")

    def uint(self) -> int:
        t = self.byte()
        if t <= 0x7F:
            return t
        if t == 0xCC:
            return self._be(">B", 1)
        if t == 0xCD:
            return self._be(">H", 2)
        if t == 0xCE:
            return self._be(">I", 4)
        if t == 0xCF:
            return self._be(">Q", 8)
        raise ValueError(f"msgpack: expected an unsigned integer at byte {self.p - 1}, found 0x{t:02x}")

    def bin8(self) -> np.ndarray:
        if self.byte() != 0xC4:
            raise ValueError(f"msgpack: expected bin8 at byte {self.p - 1}")
        n = self.byte()
        v = self.a[self.p:self.p + n]
        if v.shape[0] != n:
            raise ValueError("msgpack: truncated bin8")
        self.p += n
        return v

    def bin8_records(self, count: int, prefix: bytes = b"") -> np.ndarray:
        """`count` consecutive records `prefix ‖ c4 LEN payload`, all with the same LEN -> uint8 [count, LEN] view"""
        if count == 0:
            return np.zeros((0, 0), dtype=np.uint8)
        k = len(prefix)
        if self.p + k + 2 > self.a.shape[0] or self.a[self.p + k] != 0xC4:
            raise ValueError(f"msgpack: expected bin8 records at byte {self.p}")
        ln = int(self.a[self.p + k + 1])
        stride = k + 2 + ln
        if count > (self.a.shape[0] - self.p) // stride:
            raise ValueError("msgpack: truncated point array")
        block = self.a[self.p:self.p + count * stride]
        if block.shape[0] != count * stride:
            raise ValueError("msgpack: truncated point array")
        block = block.reshape(count, stride)
        head = np.frombuffer(prefix + bytes([0xC4, ln]), dtype=np.uint8)
        if not np.array_equal(block[:, :k + 2], np.broadcast_to(head, (count, k + 2))):
            raise ValueError("msgpack: point records of unequal size")
        self.p += count * stride
        return block[:, k + 2:]


def read_srs(path: str) -> SrsFile:
    with open(path, "rb") as f:
        buf = f.read()
    try:
        return _parse(path, buf)
    except (IndexError, struct.error, OverflowError, MemoryError) as e:   # ran off the end / absurd length fields
        raise ValueError(f"{path}: truncated or malformed msgpack ({type(e).__name__})") from e


def _parse(path: str, buf: bytes) -> SrsFile:
    r = _Reader(buf)
    fields = r.array_len()
    if fields not in (2, 3):
        raise ValueError(f"{path}: expected [g, h] or [g, h, lagrange_bases], found an array of {fields}")
    g = r.bin8_records(r.array_len())
    h = r.bin8()
    if g.shape[0] and g.shape[1] not in (33, 65) or h.shape[0] != g.shape[1]:
        raise ValueError(f"{path}: points are neither 33-byte compressed nor 65-byte uncompressed")
    out = SrsFile(g=np.ascontiguousarray(g), h=np.ascontiguousarray(h))
    if fields == 3:
        for _ in range(r.map_len()):
            n = r.uint()
            cnt = r.array_len()
            if cnt != n:
                raise ValueError(f"{path}: lagrange_bases[{n}] has {cnt} entries")
            # every entry is PolyComm{chunks}: an array of `chunks` points; all entries of one basis have the same chunk count
            save = r.p
            chunks = r.array_len()
            r.p = save
            if chunks == 1:
                pts = r.bin8_records(n, prefix=b"\x91")
                out.lagrange_bases[n] = np.ascontiguousarray(pts).reshape(n, 1, -1)
            else:
                rows = []
                for _ in range(n):
                    if r.array_len() != chunks:
                        raise ValueError(f"{path}: ragged PolyComm in lagrange_bases[{n}]")
                    rows.append(r.bin8_records(chunks))
                out.lagrange_bases[n] = np.stack(rows)
    if r.p != len(buf):
        raise ValueError(f"{path}: {len(buf) - r.p} trailing bytes")
    return out


def _w_array(n: int) -> bytes:
    return bytes([0x90 | n]) if n < 16 else (b"\xdc" + struct.pack(">H", n) if n < 65536 else b"\xdd" + struct.pack(">I", n))


def _w_uint(n: int) -> bytes:
    if n < 128:
        return bytes([n])
    if n < 256:
        return b"\xcc" + bytes([n])
    if n < 65536:
        return b"\xcd" + struct.pack(">H", n)
    return b"\xce" + struct.pack(">I", n)


def _w_points(pts: np.ndarray, prefix: bytes = b"") -> bytes:
    pts = np.ascontiguousarray(pts, dtype=np.uint8)
    n, ln = pts.shape
    k = len(prefix)
    rec = np.empty((n, k + 2 + ln), dtype=np.uint8)
    rec[:, :k + 2] = np.frombuffer(prefix + bytes([0xC4, ln]), dtype=np.uint8)
    rec[:, k + 2:] = pts
    return rec.tobytes()


def write_srs(path: str, srs: SrsFile) -> None:
    """Byte-compatible with rmp-serde's output for the same value (map order = insertion order of `lagrange_bases`)."""
    three = bool(srs.lagrange_bases) or not srs.compressed
    parts = [_w_array(3 if three else 2), _w_array(srs.g.shape[0]), _w_points(srs.g), _w_points(srs.h.reshape(1, -1))]
    if three:
        m = len(srs.lagrange_bases)
        parts.append(bytes([0x80 | m]) if m < 16 else b"\xde" + struct.pack(">H", m))
        for n, basis in srs.lagrange_bases.items():
            basis = np.asarray(basis, dtype=np.uint8)
            parts += [_w_uint(n), _w_array(n)]
            if basis.shape[1] == 1:
                parts.append(_w_points(basis[:, 0], prefix=b"\x91"))
            else:
                for row in basis:
                    parts += [_w_array(basis.shape[1]), _w_points(row)]
    with open(path, "wb") as f:
        f.write(b"".join(parts))
