"""TensorFlow checkpoints ("tensor bundles": `<prefix>.index` + `<prefix>.data-00000-of-0000N`) without TensorFlow --
the files `tf2/run.py:308-337` restores from (`tf.train.Checkpoint(model=..., global_step=..., optimizer=...)`) and the
released SimCLR checkpoints are stored in.

Format, restated from the TensorFlow sources (nothing is copied).  TensorFlow is not installable here, so no file
TensorFlow wrote has been read: what IS pinned against independent code (tests/test_host.py, using the protos TensorBoard
vendors and its CRC-checking record reader) is the `TrackableObjectGraph` / `TensorShapeProto` field numbering, the dtype
enum and the masked CRC-32C; the table layout and `BundleEntryProto` are exercised against the writer below only.

* `<prefix>.index` is an immutable sorted string table in LevelDB's table format (tensorflow/core/lib/io/table*.cc,
  format.cc): data blocks of prefix-compressed entries `varint shared | varint unshared | varint value_len | key delta |
  value`, a restart array `uint32[n] | uint32 n` at the end of every block, a 5-byte trailer per block (compression
  type, masked CRC-32C over block + type), an index block mapping separator keys to block handles, a metaindex block,
  and a 48-byte footer `metaindex handle | index handle | padding | magic 0xdb4775248b80fb57`.  Bundles are written
  uncompressed (tensor_bundle.cc); a snappy block raises.
* Key "" holds a `BundleHeaderProto{num_shards=1, endianness=2, version=3}`; every other key a
  `BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32), slices=7}`
  (tensorflow/core/protobuf/tensor_bundle.proto).  Tensor bytes are little-endian, row-major, at [offset, offset+size)
  of shard `shard_id`.
* Object-based checkpoints (TF2) store variables under keys such as `model/.../kernel/.ATTRIBUTES/VARIABLE_VALUE` and a
  serialized `TrackableObjectGraph` under `_CHECKPOINTABLE_OBJECT_GRAPH` (a scalar DT_STRING tensor: varint length,
  4-byte checksum of the lengths, bytes), whose `SerializedTensor{name=1, full_name=2, checkpoint_key=3}` attributes map
  each key to the variable's name (tensorflow/core/protobuf/trackable_object_graph.proto).  This repo names its
  variables like the reference's Keras model does, so `full_name + ':0'` is the join key.
"""
import collections
import os
import struct

import numpy as np

from .metrics import _masked_crc
from .tfrecord import _fields, _read_varint, _varint, _ld, _signed64

TABLE_MAGIC = 0xdb4775248b80fb57
OBJECT_GRAPH_KEY = '_CHECKPOINTABLE_OBJECT_GRAPH'
VARIABLE_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
           6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'), 14: 'bfloat16', 17: np.dtype('<u2'),
           19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
DT_STRING = 7
_DTYPE_CODE = {np.dtype('float32'): 1, np.dtype('float64'): 2, np.dtype('int32'): 3, np.dtype('uint8'): 4,
               np.dtype('int64'): 9, np.dtype('bool'): 10, np.dtype('float16'): 19}

Entry = collections.namedtuple('Entry', 'dtype shape shard_id offset size crc32c')


class CheckpointFormatError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------
# table (.index) reading
# ---------------------------------------------------------------------------------------------------------
def _read_handle(buf, pos):
    off, pos = _read_varint(buf, pos)
    size, pos = _read_varint(buf, pos)
    return (off, size), pos


def _read_block(data, handle, verify=True):
    off, size = handle
    if off + size + 5 > len(data):
        raise CheckpointFormatError('block handle beyond the end of the file')
    block, ctype = data[off:off + size], data[off + size]
    if verify:
        (crc,) = struct.unpack('<I', data[off + size + 1:off + size + 5])
        if _masked_crc(data[off:off + size + 1]) != crc:
            raise CheckpointFormatError('block checksum mismatch')
    if ctype != 0:
        raise CheckpointFormatError('compressed table block (type %d): tensor bundles are written uncompressed' % ctype)
    return block


def _block_entries(block):
    """(key bytes, value bytes) of a table block, undoing the shared-prefix compression."""
    if len(block) < 4:
        raise CheckpointFormatError('table block too short')
    (num_restarts,) = struct.unpack('<I', block[-4:])
    limit = len(block) - 4 * (num_restarts + 1)
    if limit < 0:
        raise CheckpointFormatError('bad restart array')
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _read_varint(block, pos)
        unshared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > limit:
            raise CheckpointFormatError('corrupt table entry')
        key = key[:shared] + bytes(block[pos:pos + unshared]); pos += unshared
        yield key, bytes(block[pos:pos + vlen]); pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    data = open(path, 'rb').read()
    if len(data) < 48:
        raise CheckpointFormatError('%s: shorter than a table footer' % path)
    footer = data[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise CheckpointFormatError('%s: not a table file (bad magic)' % path)
    _, pos = _read_handle(footer, 0)                  # metaindex: unused by bundles
    index_handle, _ = _read_handle(footer, pos)
    out = []
    for _, hv in _block_entries(_read_block(data, index_handle, verify)):
        handle, _ = _read_handle(hv, 0)
        out.extend(_block_entries(_read_block(data, handle, verify)))
    return out


# ---------------------------------------------------------------------------------------------------------
# bundle reading
# ---------------------------------------------------------------------------------------------------------
def _parse_shape(buf):
    dims = []
    for num, wt, val in _fields(buf):
        if num == 2 and wt == 2:                      # Dim{size=1, name=2}
            size = 0
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 0:
                    size = _signed64(v2)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    dtype, shape, shard, off, size, crc = 0, (), 0, 0, 0, 0
    for num, wt, val in _fields(memoryview(buf)):
        if num == 1 and wt == 0: dtype = val
        elif num == 2 and wt == 2: shape = _parse_shape(val)
        elif num == 3 and wt == 0: shard = val
        elif num == 4 and wt == 0: off = val
        elif num == 5 and wt == 0: size = val
        elif num == 6 and wt == 5: (crc,) = struct.unpack('<I', val)
        elif num == 7:
            raise CheckpointFormatError('partitioned (sliced) variables are not supported')
    return Entry(dtype, shape, shard, off, size, crc)


class TensorBundleReader:
    def __init__(self, prefix, verify=True):
        self.prefix = prefix
        index = prefix + '.index'
        if not os.path.exists(index):
            raise FileNotFoundError(index)
        self.num_shards, self.entries = 1, collections.OrderedDict()
        for key, val in read_table(index, verify):
            if key == b'':
                for num, wt, v in _fields(memoryview(val)):
                    if num == 1 and wt == 0: self.num_shards = v
                    elif num == 2 and wt == 0 and v != 0:
                        raise CheckpointFormatError('big-endian bundle')
            else:
                self.entries[key.decode('utf-8')] = _parse_entry(val)
        self._shards = {}

    def keys(self):
        return list(self.entries)

    def _bytes(self, e):
        if e.shard_id not in self._shards:
            path = '%s.data-%05d-of-%05d' % (self.prefix, e.shard_id, self.num_shards)
            self._shards[e.shard_id] = np.memmap(path, dtype=np.uint8, mode='r') if os.path.getsize(path) else np.zeros(0, np.uint8)
        buf = self._shards[e.shard_id]
        if e.offset + e.size > len(buf):
            raise CheckpointFormatError('tensor beyond the end of its shard')
        return bytes(buf[e.offset:e.offset + e.size])

    def get_string(self, key):
        """Scalar DT_STRING tensor (the object graph)."""
        e = self.entries[key]
        if e.dtype != DT_STRING:
            raise CheckpointFormatError('%s is not a string tensor' % key)
        raw = self._bytes(e)
        n, pos = _read_varint(raw, 0)
        if pos + 4 + n == len(raw):
            pos += 4                                   # checksum of the length varints
        elif pos + n != len(raw):
            raise CheckpointFormatError('unexpected string tensor layout')
        return raw[pos:pos + n]

    def get_tensor(self, key):
        e = self.entries[key]
        if e.dtype == DT_STRING:
            return self.get_string(key)
        if e.dtype not in _DTYPES:
            raise CheckpointFormatError('%s: unsupported dtype %d' % (key, e.dtype))
        raw, dt = self._bytes(e), _DTYPES[e.dtype]
        if dt == 'bfloat16':
            a = (np.frombuffer(raw, dtype='<u2').astype(np.uint32) << 16).view(np.float32)
        else:
            a = np.frombuffer(raw, dtype=dt)
        n = int(np.prod(e.shape)) if e.shape else 1
        if a.size != n:
            raise CheckpointFormatError('%s: %d elements for shape %s' % (key, a.size, e.shape))
        return a.reshape(e.shape).copy()

    def object_graph(self):
        """[(checkpoint_key, full_name, attribute name)] of every serialized tensor of the TrackableObjectGraph."""
        if OBJECT_GRAPH_KEY not in self.entries:
            return []
        out = []
        for num, wt, node in _fields(memoryview(self.get_string(OBJECT_GRAPH_KEY))):
            if num != 1 or wt != 2:
                continue
            for n2, w2, attr in _fields(node):
                if n2 != 2 or w2 != 2:
                    continue
                name = full = ckey = ''
                for n3, w3, v in _fields(attr):
                    if w3 != 2: continue
                    if n3 == 1: name = bytes(v).decode('utf-8')
                    elif n3 == 2: full = bytes(v).decode('utf-8')
                    elif n3 == 3: ckey = bytes(v).decode('utf-8')
                out.append((ckey, full, name))
        return out

    def variables_by_name(self):
        """{variable name: array}.  Object-based checkpoints: `full_name` of the object graph (first key wins when a
        variable is reachable under several names); name-based (TF1) checkpoints: the keys themselves."""
        out = collections.OrderedDict()
        graph = self.object_graph()
        if graph:
            for ckey, full, name in graph:
                if name == 'VARIABLE_VALUE' and ckey in self.entries and full and full not in out:
                    out[full] = self.get_tensor(ckey)
        else:
            for k, e in self.entries.items():
                if e.dtype != DT_STRING:
                    out[k] = self.get_tensor(k)
        return out


def is_tf_checkpoint(path):
    return bool(path) and os.path.exists(path + '.index')


# ---------------------------------------------------------------------------------------------------------
# writing (tests; exporting weights for TensorFlow-side tools)
# ---------------------------------------------------------------------------------------------------------
def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_table(path, items, block_entries=64):
    """items: sorted [(key bytes, value bytes)]."""
    data, index = bytearray(), []

    def emit(block):
        off = len(data)
        data.extend(block + b'\x00')
        data.extend(struct.pack('<I', _masked_crc(block + b'\x00')))
        return _varint(off) + _varint(len(block))
    for i in range(0, max(len(items), 1), block_entries):
        chunk = items[i:i + block_entries]
        if chunk:
            index.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    data.extend(footer)
    with open(path, 'wb') as fh:
        fh.write(bytes(data))


def _shape_proto(shape):
    return b''.join(_ld(2, _varint((1 << 3) | 0) + _varint(int(d))) for d in shape)


def write_bundle(prefix, tensors, names=None):
    """tensors: {checkpoint key: ndarray}; names: {checkpoint key: variable full_name} -> object graph (one node per
    variable under the root).  Single shard, little endian."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items, blob = [], bytearray()
    tensors = dict(tensors)
    if names:
        root = b''.join(_ld(1, _varint((1 << 3) | 0) + _varint(i + 1) + _ld(2, ('v%d' % i).encode())) for i in range(len(names)))
        nodes = _ld(1, root)
        for key, full in names.items():
            attr = _ld(1, b'VARIABLE_VALUE') + _ld(2, full.encode()) + _ld(3, key.encode())
            nodes += _ld(1, _ld(2, attr))
        tensors[OBJECT_GRAPH_KEY] = nodes
    for key in sorted(tensors, key=lambda k: k.encode('utf-8')):
        v = tensors[key]
        off = len(blob)
        if isinstance(v, (bytes, bytearray)):
            lens = _varint(len(v))
            raw = lens + struct.pack('<I', _masked_crc(lens)) + bytes(v)
            dtype, shape = DT_STRING, ()
        else:
            a = np.asarray(v)                       # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in _DTYPE_CODE:
                raise CheckpointFormatError('cannot write dtype %s' % a.dtype)
            raw, dtype, shape = a.astype(a.dtype.newbyteorder('<')).tobytes(order='C'), _DTYPE_CODE[a.dtype], a.shape
        blob += raw
        entry = _varint((1 << 3) | 0) + _varint(dtype) + _ld(2, _shape_proto(shape))
        entry += _varint((4 << 3) | 0) + _varint(off) + _varint((5 << 3) | 0) + _varint(len(raw))
        entry += _varint((6 << 3) | 5) + struct.pack('<I', _masked_crc(raw))
        items.append((key.encode('utf-8'), entry))
    header = _varint((1 << 3) | 0) + _varint(1) + _ld(3, _varint((1 << 3) | 0) + _varint(1))     # num_shards 1, version.producer 1
    write_table(prefix + '.index', [(b'', header)] + items)
    with open(prefix + '.data-00000-of-00001', 'wb') as fh:
        fh.write(bytes(blob))


def main(argv=None):
    """`python -m simclr_b200.tf_checkpoint <prefix>`: list the tensors of a TensorFlow checkpoint (key, dtype, shape)
    and the variable name each object-graph key maps to."""
    import sys
    args = list(sys.argv[1:] if argv is None else argv)
    if len(args) != 1:
        print('usage: python -m simclr_b200.tf_checkpoint <checkpoint prefix>')
        return 2
    r = TensorBundleReader(args[0])
    names = {k: full for k, full, attr in r.object_graph() if attr == 'VARIABLE_VALUE'}
    for k, e in r.entries.items():
        dt = 'string' if e.dtype == DT_STRING else str(_DTYPES.get(e.dtype, 'dtype %d' % e.dtype))
        print('%-90s %-9s %-20s %s' % (k, dt, list(e.shape), names.get(k, '')))
    print('%d tensors, %d shard(s)' % (len(r.entries), r.num_shards))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
