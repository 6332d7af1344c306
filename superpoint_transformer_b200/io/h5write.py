"""Write side of io/h5lite.py: the HDF5 subset the reference's `NAG.save` produces through h5py
(superblock 0, version-1 object headers, old-style groups = B-tree + local heap + symbol-table
nodes, contiguous little-endian integer / float datasets, one scalar attribute per group,
variable-length string lists through a global heap collection).  The structures are laid out
like the ones in the reference's own file (notebooks/demo_nag_v3.h5: same message types, versions
and field values); verified by reading the result back with io/h5lite.py — there is no libhdf5
in this environment to give a second opinion.

    write_h5(path, tree, attrs)      tree: {name: numpy array | list of str | nested dict}
"""
import numpy as np

__all__ = ['write_h5']

_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _NODE_K = 4, 16            # group B-tree ranks stored in the superblock (h5py defaults)


def _u(value, n):
    return int(value).to_bytes(n, 'little')


def _pad8(b):
    return b + bytes((-len(b)) % 8)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return _u(mtype, 2) + _u(len(body), 2) + _u(flags, 1) + bytes(3) + body


def _object_header(messages):
    data = b''.join(messages)
    return (bytes([1, 0]) + _u(len(messages), 2) + _u(1, 4) + _u(len(data), 4) + bytes(4)
            + data)


def _datatype(dtype):
    dtype = np.dtype(dtype)
    if dtype.byteorder == '>':
        raise NotImplementedError('big-endian arrays')
    size = dtype.itemsize
    if dtype.kind in 'iub':
        signed = 0x08 if dtype.kind == 'i' else 0x00
        return bytes([0x10, signed, 0, 0]) + _u(size, 4) + _u(0, 2) + _u(8 * size, 2)
    if dtype.kind == 'f':
        # (exponent location, exponent size, mantissa size, bias) of IEEE half / single / double
        exp_loc, exp_size, man_size, bias = {2: (10, 5, 10, 15), 4: (23, 8, 23, 127),
                                             8: (52, 11, 52, 1023)}[size]
        return (bytes([0x11, 0x20, 8 * size - 1, 0]) + _u(size, 4) + _u(0, 2) + _u(8 * size, 2)
                + bytes([exp_loc, exp_size, 0, man_size]) + _u(bias, 4))
    raise NotImplementedError(f'dtype {dtype}')


_VLEN_STR_TYPE = bytes([0x19, 0x01, 0x01, 0x00]) + _u(16, 4) + \
    bytes([0x10, 0, 0, 0]) + _u(1, 4) + _u(0, 2) + _u(8, 2)


def _dataspace(shape):
    dims = b''.join(_u(d, 8) for d in shape)
    return bytes([1, len(shape), 1, 0]) + bytes(4) + dims + dims      # (max dims = dims)


class _File:
    def __init__(self):
        self.buf = bytearray(96)            # superblock + root symbol-table entry, filled last

    def put(self, data):
        self.buf += bytes((-len(self.buf)) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # -- datasets ------------------------------------------------------------------------
    def dataset(self, array):
        array = np.ascontiguousarray(array)
        if array.dtype == np.bool_:
            array = array.astype(np.uint8)
        raw = array.tobytes()
        addr = self.put(raw) if raw else _UNDEF
        return self._dataset_header(_datatype(array.dtype), array.shape, addr, len(raw))

    def strings(self, items):
        """list of str -> variable-length string dataset backed by one global heap collection"""
        if not items:
            return self._dataset_header(_datatype(np.float64), (0,), _UNDEF, 0)
        objects = b''
        for i, s in enumerate(items):
            data = s.encode('utf-8')
            objects += _u(i + 1, 2) + _u(1, 2) + bytes(4) + _u(len(data), 8) + _pad8(data)
        size = max(4096, 16 + len(objects) + 16)
        free = size - 16 - len(objects)
        heap = b'GCOL' + bytes([1, 0, 0, 0]) + _u(size, 8) + objects
        heap += _u(0, 2) + _u(0, 2) + bytes(4) + _u(free, 8) + bytes(free - 16)
        heap_addr = self.put(heap)
        raw = b''.join(_u(len(s.encode('utf-8')), 4) + _u(heap_addr, 8) + _u(i + 1, 4)
                       for i, s in enumerate(items))
        return self._dataset_header(_VLEN_STR_TYPE, (len(items),), self.put(raw), len(raw))

    def _dataset_header(self, datatype, shape, addr, nbytes):
        fill_time = 0 if datatype is _VLEN_STR_TYPE else 2      # as h5py writes them
        return self.put(_object_header([
            _message(0x01, _dataspace(shape)),
            _message(0x03, datatype, flags=1),
            _message(0x05, bytes([2, 2, fill_time, 1]) + _u(0, 4), flags=1),   # no fill value
            _message(0x08, bytes([3, 1]) + _u(addr, 8) + _u(nbytes, 8)),    # contiguous layout
        ]))

    # -- groups --------------------------------------------------------------------------
    def group(self, tree, attrs=None):
        """Returns (object header, B-tree, local heap) addresses."""
        entries = []
        for name in sorted(tree, key=lambda s: s.encode('utf-8')):
            value = tree[name]
            if isinstance(value, dict):
                entries.append((name,) + self.group(value))
            elif isinstance(value, (list, tuple)) and all(isinstance(s, str) for s in value):
                entries.append((name, self.strings(list(value)), None, None))
            else:
                entries.append((name, self.dataset(value), None, None))
        # local heap: the empty string at offset 0, then the names, 8-byte aligned
        heap_data, offsets = bytearray(8), []
        for name, *_ in entries:
            offsets.append(len(heap_data))
            heap_data += _pad8(name.encode('utf-8') + b'\0')
        data_addr = self.put(bytes(heap_data))
        heap_addr = self.put(b'HEAP' + bytes(4) + _u(len(heap_data), 8) + _u(1, 8)
                             + _u(data_addr, 8))
        # symbol-table nodes of <= 2K entries, children of one B-tree leaf
        per_node = 2 * _LEAF_K
        if len(entries) > per_node * 2 * _NODE_K:
            raise NotImplementedError('groups with more than 256 members')
        children, keys = [], [0]
        for start in range(0, len(entries), per_node):
            chunk = list(zip(entries[start:start + per_node], offsets[start:start + per_node]))
            node = b'SNOD' + bytes([1, 0]) + _u(len(chunk), 2)
            for (name, header, btree, heap), off in chunk:
                scratch = bytes(16) if btree is None else _u(btree, 8) + _u(heap, 8)
                node += _u(off, 8) + _u(header, 8) + _u(0 if btree is None else 1, 4) \
                    + bytes(4) + scratch
            node += bytes(40 * (per_node - len(chunk)))
            children.append(self.put(node))
            keys.append(chunk[-1][1])
        tree_node = b'TREE' + bytes([0, 0]) + _u(len(children), 2) + _u(_UNDEF, 8) + _u(_UNDEF, 8)
        for i, child in enumerate(children):
            tree_node += _u(keys[i], 8) + _u(child, 8)
        tree_node += _u(keys[len(children)], 8)
        tree_node += bytes(24 + (2 * _NODE_K + 1) * 8 + 2 * _NODE_K * 8 - len(tree_node))
        btree_addr = self.put(tree_node)
        messages = [_message(0x11, _u(btree_addr, 8) + _u(heap_addr, 8))]
        for key, value in (attrs or {}).items():
            name = key.encode('utf-8') + b'\0'
            dtype = _datatype(np.int64)
            space = bytes([1, 0, 0, 0]) + bytes(4)
            messages.append(_message(0x0C, bytes([1, 0]) + _u(len(name), 2) + _u(len(dtype), 2)
                                     + _u(len(space), 2) + _pad8(name) + _pad8(dtype)
                                     + _pad8(space) + _u(int(value) & _UNDEF, 8)))
        return self.put(_object_header(messages)), btree_addr, heap_addr

    def finish(self, root):
        header, btree, heap = root
        self.buf += bytes((-len(self.buf)) % 8)
        sb = b'\x89HDF\r\n\x1a\n' + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + _u(_LEAF_K, 2) \
            + _u(_NODE_K, 2) + _u(0, 4) + _u(0, 8) + _u(_UNDEF, 8) + _u(len(self.buf), 8) \
            + _u(_UNDEF, 8)
        sb += _u(0, 8) + _u(header, 8) + _u(1, 4) + bytes(4) + _u(btree, 8) + _u(heap, 8)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, tree, attrs=None):
    """Write `tree` ({name: numpy array | list of str | nested dict}) as an HDF5 file with the
    integer root attributes `attrs`."""
    f = _File()
    data = f.finish(f.group(tree, attrs))
    with open(path, 'wb') as fh:
        fh.write(data)
