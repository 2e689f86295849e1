"""
TFRecord files of serialised `tf.train.Example` messages, without TensorFlow.

The reference stores every structure of a dataset as one `tf.train.Example` whose features are
raw byte strings (`transformer/base.py:229-247,365-437`, `transformer/universal.py:1177-1330`)
and writes them with `tf.io.TFRecordWriter` (`train/dataset/dataset.py:168-258`).  Both layers
are public, stable formats; this module restates them:

* TFRecord framing (tensorflow/core/lib/io/record_writer.cc): per record
  `uint64 length | uint32 masked_crc32c(length) | bytes data | uint32 masked_crc32c(data)`,
  little endian, `masked(c) = rotr(c, 15) + 0xa282ead8 (mod 2^32)`, CRC-32C (Castagnoli,
  reflected polynomial 0x82F63B78).
* `Example` protobuf (tensorflow/core/example/{example,feature}.proto):
      Example   { Features features = 1; }
      Features  { map<string, Feature> feature = 1; }        (entry: key = 1, value = 2)
      Feature   { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2;
                               Int64List int64_list = 3; } }
      BytesList { repeated bytes value = 1; }
      FloatList { repeated float value = 1 [packed = true]; }
      Int64List { repeated int64 value = 1 [packed = true]; }

`tests/test_tfrecord.py` checks the message codec against google.protobuf (the same schema
built from descriptors) in both directions, and the CRC against its published check value.
"""
import struct
from typing import Dict, Iterable, Iterator, List, Union

import numpy as np

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.arange(256, dtype=np.uint32)
        for _ in range(8):
            t = np.where(t & 1, (t >> 1) ^ np.uint32(0x82F63B78), t >> 1).astype(np.uint32)
        # slicing-by-8 tables: T[k][b] = crc of byte b followed by k zero bytes
        tabs = [t]
        for _ in range(7):
            prev = tabs[-1]
            tabs.append((prev >> 8) ^ t[prev & 0xFF])
        _CRC_TABLE = [[int(v) for v in tab] for tab in tabs]
    return _CRC_TABLE


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C of `data` (check value: crc32c(b'123456789') = 0xE3069283)."""
    T = _crc_table()
    t0, t1, t2, t3, t4, t5, t6, t7 = T
    crc ^= 0xFFFFFFFF
    n8 = len(data) // 8
    if n8:
        for lo, hi in struct.iter_unpack('<II', memoryview(data)[:8 * n8]):
            lo ^= crc
            crc = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
                   t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in data[8 * n8:]:
        crc = t0[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# -- protobuf wire format -----------------------------------------------------------------
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64            # int64 two's complement, ten bytes
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _ld(field: int, payload: bytes) -> bytes:
    """length-delimited field"""
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf) -> Iterator:
    """(field number, wire type, value) of every field of a message; value = int for varint /
    fixed types, memoryview for length-delimited ones."""
    buf = memoryview(buf)
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            if pos + n > end:
                raise ValueError("truncated message")
            val = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            val = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield field, wt, val


FeatureValue = Union[List[bytes], List[float], List[int]]


class Feature:
    """One `tf.train.Feature`: `kind` in {'bytes_list', 'float_list', 'int64_list'}."""
    __slots__ = ('kind', 'value')

    def __init__(self, kind: str, value: FeatureValue):
        assert kind in ('bytes_list', 'float_list', 'int64_list')
        self.kind, self.value = kind, list(value)

    def __eq__(self, other):
        return isinstance(other, Feature) and self.kind == other.kind and \
            self.value == other.value

    def __repr__(self):
        return f"Feature({self.kind}, n={len(self.value)})"

    def SerializeToString(self) -> bytes:
        if self.kind == 'bytes_list':
            return _ld(1, b''.join(_ld(1, bytes(v)) for v in self.value))
        if self.kind == 'float_list':
            packed = struct.pack(f'<{len(self.value)}f', *self.value)
            return _ld(2, _ld(1, packed) if self.value else b'')
        packed = b''.join(_varint(int(v)) for v in self.value)
        return _ld(3, _ld(1, packed) if self.value else b'')

    @classmethod
    def FromString(cls, buf) -> 'Feature':
        out = None
        for field, wt, val in _fields(buf):
            if wt != 2 or field not in (1, 2, 3):
                continue
            if field == 1:
                out = cls('bytes_list', [bytes(v) for f, w, v in _fields(val)
                                         if f == 1 and w == 2])
            elif field == 2:
                vals = []
                for f, w, v in _fields(val):
                    if f != 1:
                        continue
                    if w == 2:                                  # packed
                        vals.extend(struct.unpack(f'<{len(v) // 4}f', v))
                    elif w == 5:
                        vals.append(struct.unpack('<f', struct.pack('<I', v))[0])
                out = cls('float_list', vals)
            else:
                vals = []
                for f, w, v in _fields(val):
                    if f != 1:
                        continue
                    if w == 2:
                        p = 0
                        while p < len(v):
                            x, p = _read_varint(v, p)
                            vals.append(x - (1 << 64) if x >> 63 else x)
                    elif w == 0:
                        vals.append(v - (1 << 64) if v >> 63 else v)
                out = cls('int64_list', vals)
        if out is None:
            out = cls('bytes_list', [])
        return out


def bytes_feature(value: bytes) -> Feature:
    """transformer/base.py:229-233"""
    return Feature('bytes_list', [value])


def float_feature(value: float) -> Feature:
    """transformer/base.py:236-240"""
    return Feature('float_list', [value])


def int64_feature(value: int) -> Feature:
    """transformer/base.py:243-247"""
    return Feature('int64_list', [value])


class Example:
    """`tf.train.Example` restricted to what the reference uses: a flat feature map."""

    def __init__(self, features: Dict[str, Feature] = None):
        self.features = dict(features or {})

    def __eq__(self, other):
        return isinstance(other, Example) and self.features == other.features

    def SerializeToString(self) -> bytes:
        """Deterministic: map entries in key order (protobuf leaves the order open)."""
        entries = b''.join(
            _ld(1, _ld(1, k.encode('utf-8')) + _ld(2, self.features[k].SerializeToString()))
            for k in sorted(self.features))
        return _ld(1, entries)

    @classmethod
    def FromString(cls, buf) -> 'Example':
        feats = {}
        for field, wt, val in _fields(buf):
            if field != 1 or wt != 2:
                continue
            for f2, w2, entry in _fields(val):
                if f2 != 1 or w2 != 2:
                    continue
                key, feat = '', Feature('bytes_list', [])
                for f3, w3, v3 in _fields(entry):
                    if f3 == 1 and w3 == 2:
                        key = bytes(v3).decode('utf-8')
                    elif f3 == 2 and w3 == 2:
                        feat = Feature.FromString(v3)
                feats[key] = feat
        return cls(feats)


# -- record files -------------------------------------------------------------------------
class TFRecordWriter:
    """`tf.io.TFRecordWriter(filename)` (uncompressed): `write(bytes)`, context manager."""

    def __init__(self, filename: str):
        self._fp = open(filename, 'wb')

    def write(self, record: bytes):
        head = struct.pack('<Q', len(record))
        self._fp.write(head)
        self._fp.write(struct.pack('<I', masked_crc32c(head)))
        self._fp.write(record)
        self._fp.write(struct.pack('<I', masked_crc32c(record)))

    def flush(self):
        self._fp.flush()

    def close(self):
        self._fp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def read_tfrecords(filename: str, verify: bool = True) -> Iterator[bytes]:
    """The records of one file (what `tf.data.TFRecordDataset([filename])` yields).  A
    checksum mismatch or a truncated record raises `IOError` (TensorFlow: DataLossError)."""
    with open(filename, 'rb') as fp:
        while True:
            head = fp.read(8)
            if not head:
                return
            if len(head) < 8:
                raise IOError(f"{filename}: truncated record header")
            crc = fp.read(4)
            if len(crc) < 4:
                raise IOError(f"{filename}: truncated record header")
            if verify and struct.unpack('<I', crc)[0] != masked_crc32c(head):
                raise IOError(f"{filename}: corrupted record length")
            n = struct.unpack('<Q', head)[0]
            data = fp.read(n)
            crc = fp.read(4)
            if len(data) < n or len(crc) < 4:
                raise IOError(f"{filename}: truncated record")
            if verify and struct.unpack('<I', crc)[0] != masked_crc32c(data):
                raise IOError(f"{filename}: corrupted record")
            yield data


def write_tfrecords(filename: str, records: Iterable[bytes]) -> int:
    n = 0
    with TFRecordWriter(filename) as w:
        for r in records:
            w.write(r)
            n += 1
    return n
