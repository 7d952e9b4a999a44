"""TFRecord files and `tf.train.Example` records without TensorFlow: the on-disk format of the TFDS datasets the
reference reads (`tf2/run.py:469-475` `tfds.builder(...)`, `tf2/data.py:64-73` `builder.as_dataset`).

* TFRecord framing (tensorflow/core/lib/io/record_writer.cc): `uint64 length | uint32 masked_crc32c(length) |
  data | uint32 masked_crc32c(data)`, little endian -- the framing `metrics.SummaryWriter` already writes.
* `tf.train.Example` (tensorflow/core/example/{example,feature}.proto): `Example{features=1}`,
  `Features{map<string, Feature> feature=1}`, `Feature{oneof: BytesList bytes_list=1 | FloatList float_list=2 |
  Int64List int64_list=3}`, each list a repeated field 1 (floats / int64s usually packed).  Parsed by hand: four
  nested length-delimited messages, no generated code.

The writer half exists for the tests and for converting other sources into the layout `data.TFRecordBuilder` reads.
"""
import struct

from .metrics import _masked_crc


class TFRecordError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------
# record framing
# ---------------------------------------------------------------------------------------------------------
def read_records(path, verify_data_crc=False):
    """Yields the payload of every record of one file.  The 12-byte header is always CRC-checked (it guards the
    length the reader is about to trust); the payload CRC is a pure-Python byte loop and therefore optional."""
    with open(path, 'rb') as fh:
        while True:
            head = fh.read(12)
            if not head:
                return
            if len(head) < 12:
                raise TFRecordError('%s: truncated record header' % path)
            (length,), (crc,) = struct.unpack('<Q', head[:8]), struct.unpack('<I', head[8:])
            if _masked_crc(head[:8]) != crc:
                raise TFRecordError('%s: corrupt record header' % path)
            data = fh.read(length)
            tail = fh.read(4)
            if len(data) < length or len(tail) < 4:
                raise TFRecordError('%s: truncated record' % path)
            if verify_data_crc and _masked_crc(data) != struct.unpack('<I', tail)[0]:
                raise TFRecordError('%s: corrupt record payload' % path)
            yield data


def count_records(path):
    """Number of records, reading only the headers."""
    n = 0
    with open(path, 'rb') as fh:
        while True:
            head = fh.read(12)
            if len(head) < 12:
                return n
            fh.seek(struct.unpack('<Q', head[:8])[0] + 4, 1)
            n += 1


def write_records(path, payloads):
    with open(path, 'wb') as fh:
        for data in payloads:
            head = struct.pack('<Q', len(data))
            fh.write(head + struct.pack('<I', _masked_crc(head)) + data + struct.pack('<I', _masked_crc(data)))


# ---------------------------------------------------------------------------------------------------------
# protobuf wire format (the subset Example uses)
# ---------------------------------------------------------------------------------------------------------
def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise TFRecordError('truncated varint')
        b = buf[pos]; pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise TFRecordError('varint too long')


def _fields(buf):
    """(field number, wire type, value) of one message; value: int (varint / fixed) or memoryview (length-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val = buf[pos:pos + ln]; pos += ln
            if len(val) < ln:
                raise TFRecordError('truncated field')
        elif wt == 5:
            val = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise TFRecordError('unsupported wire type %d' % wt)
        yield num, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf):
    for num, wt, val in _fields(buf):
        if wt != 2:
            continue
        if num == 1:        # BytesList
            return [bytes(v) for n, w, v in _fields(val) if n == 1 and w == 2]
        if num == 2:        # FloatList: packed (wire type 2) or one fixed32 per element
            out = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    out.extend(struct.unpack('<%df' % (len(v) // 4), bytes(v)))
                elif n == 1 and w == 5:
                    out.append(struct.unpack('<f', v)[0])
            return out
        if num == 3:        # Int64List: packed varints or one varint per element
            out = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    p = 0
                    while p < len(v):
                        x, p = _read_varint(v, p)
                        out.append(_signed64(x))
                elif n == 1 and w == 0:
                    out.append(_signed64(v))
            return out
    return []


def parse_example(data):
    """Serialized `tf.train.Example` -> {feature name: list of bytes / float / int}."""
    out = {}
    buf = memoryview(data)
    for num, wt, features in _fields(buf):
        if num != 1 or wt != 2:
            continue
        for n2, w2, entry in _fields(features):            # map<string, Feature> entries
            if n2 != 1 or w2 != 2:
                continue
            key, feat = None, None
            for n3, w3, v in _fields(entry):
                if n3 == 1 and w3 == 2:
                    key = bytes(v).decode('utf-8')
                elif n3 == 2 and w3 == 2:
                    feat = v
            if key is not None:
                out[key] = _parse_feature(feat) if feat is not None else []
    return out


def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num, payload):
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features):
    """{name: bytes | [bytes] | int | [int] | float | [float]} -> serialized `tf.train.Example` (packed lists)."""
    entries = b''
    for name in sorted(features):
        v = features[name]
        if isinstance(v, (bytes, bytearray, int, float)):
            v = [v]
        if all(isinstance(x, (bytes, bytearray)) for x in v):
            feat = _ld(1, b''.join(_ld(1, bytes(x)) for x in v))
        elif all(isinstance(x, int) for x in v):
            feat = _ld(3, _ld(1, b''.join(_varint(x) for x in v)))
        else:
            feat = _ld(2, _ld(1, struct.pack('<%df' % len(v), *[float(x) for x in v])))
        entries += _ld(1, _ld(1, name.encode('utf-8')) + _ld(2, feat))
    return _ld(1, entries)
