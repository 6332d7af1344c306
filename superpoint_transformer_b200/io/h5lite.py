"""Read-only HDF5 access for the files the reference writes with h5py (`NAG.save`, reference
src/data/nag.py:401-470, src/utils/io.py:16-67): no h5py in this environment, so the subset of
the format those files use is parsed directly —

    superblock version 0, version-1 object headers (with continuation blocks), old-style groups
    (version-1 B-tree + local heap + symbol-table nodes), dataspace v1/v2, little-endian
    fixed-point / IEEE float datatypes, contiguous or compact layout (message version 3),
    scalar / 1-D numeric attributes, variable-length strings through the global heap.

Anything else (chunked or filtered datasets, new-style groups, compound types, big-endian data)
raises NotImplementedError naming the feature: the reference saves plain uncompressed datasets
(`save_tensor` calls `create_dataset(key, data=...)` without chunking).  Host-side I/O: numpy
arrays come back, nothing here touches the GPU.
"""
import numpy as np

__all__ = ['H5File', 'H5Group', 'H5Dataset']

_UNDEF = 0xFFFFFFFFFFFFFFFF
_MSG_DATASPACE, _MSG_DATATYPE, _MSG_LAYOUT, _MSG_FILTERS = 0x1, 0x3, 0x8, 0xB
_MSG_ATTRIBUTE, _MSG_CONTINUATION, _MSG_SYMBOL_TABLE = 0xC, 0x10, 0x11


def _pad8(n):
    return (n + 7) & ~7


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        if buf[:8] != b'\x89HDF\r\n\x1a\n':
            raise ValueError('not an HDF5 file')
        if buf[8] != 0:
            raise NotImplementedError(f'HDF5 superblock version {buf[8]} (only 0 is supported)')
        self.O, self.L = buf[13], buf[14]
        if (self.O, self.L) != (8, 8):
            raise NotImplementedError('HDF5 files with offsets / lengths that are not 8 bytes')
        self.base = self.u(24, 8)
        self.root_entry = 24 + 4 * self.O

    def u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], 'little')

    def cstr(self, off):
        end = self.buf.index(b'\0', off)
        return self.buf[off:end].decode('utf-8')

    # -- object headers ------------------------------------------------------------------
    def messages(self, addr):
        """[(type, size, body offset)] of a version-1 object header, continuations followed."""
        if self.buf[addr] != 1:
            raise NotImplementedError(f'HDF5 object header version {self.buf[addr]}')
        count, size = self.u(addr + 2, 2), self.u(addr + 8, 4)
        blocks, out = [(addr + 16, size)], []
        while blocks:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < count:
                mtype, msize, body = self.u(p, 2), self.u(p + 2, 2), p + 8
                if mtype == _MSG_CONTINUATION:
                    blocks.append((self.u(body, 8), self.u(body + 8, 8)))
                out.append((mtype, msize, body))
                p = body + msize
        return out

    # -- groups --------------------------------------------------------------------------
    def group_links(self, btree, heap):
        """{name: object header address} of an old-style group."""
        if self.buf[heap:heap + 4] != b'HEAP':
            raise ValueError('corrupt local heap')
        heap_data = self.u(heap + 24, 8)
        links = {}

        def node(addr):
            if self.buf[addr:addr + 4] != b'TREE':
                raise ValueError('corrupt group B-tree')
            if self.buf[addr + 4] != 0:
                raise NotImplementedError('chunked datasets (B-tree node type 1)')
            level, used = self.buf[addr + 5], self.u(addr + 6, 2)
            p = addr + 8 + 2 * self.O
            for _ in range(used):
                child = self.u(p + self.L, 8)
                p += self.L + self.O
                if level > 0:
                    node(child)
                    continue
                if self.buf[child:child + 4] != b'SNOD':
                    raise ValueError('corrupt symbol table node')
                q = child + 8
                for _ in range(self.u(child + 6, 2)):
                    links[self.cstr(heap_data + self.u(q, 8))] = self.u(q + 8, 8)
                    q += 2 * self.O + 24
        node(btree)
        return links

    # -- datatypes -----------------------------------------------------------------------
    def dtype(self, body):
        """numpy dtype of a datatype message, or ('vlen_str',) for variable-length strings."""
        cls, version = self.buf[body] & 15, self.buf[body] >> 4
        bits0, size = self.buf[body + 1], self.u(body + 4, 4)
        if cls in (0, 1) and (bits0 & 1):
            raise NotImplementedError('big-endian HDF5 data')
        if cls == 0:
            return np.dtype(('<i' if bits0 & 8 else '<u') + str(size))
        if cls == 1:
            return np.dtype('<f' + str(size))
        if cls == 9 and (bits0 & 15) == 1:
            return ('vlen_str',)
        raise NotImplementedError(f'HDF5 datatype class {cls} (version {version})')

    def shape(self, body):
        version, rank = self.buf[body], self.buf[body + 1]
        if version not in (1, 2):
            raise NotImplementedError(f'HDF5 dataspace version {version}')
        start = body + (8 if version == 1 else 4)
        return tuple(self.u(start + 8 * i, 8) for i in range(rank))

    def vlen_strings(self, raw, count):
        """Variable-length strings: each element = length (4) + global heap address (8) +
        object index (4)."""
        out = []
        for i in range(count):
            p = 16 * i
            length = int.from_bytes(raw[p:p + 4], 'little')
            heap = int.from_bytes(raw[p + 4:p + 12], 'little')
            index = int.from_bytes(raw[p + 12:p + 16], 'little')
            out.append(self.global_heap_object(heap, index)[:length].decode('utf-8'))
        return out

    def global_heap_object(self, addr, index):
        if self.buf[addr:addr + 4] != b'GCOL':
            raise ValueError('corrupt global heap collection')
        end = addr + self.u(addr + 8, 8)
        p = addr + 16
        while p + 16 <= end:
            obj, size = self.u(p, 2), self.u(p + 8, 8)
            if obj == index:
                return self.buf[p + 16:p + 16 + size]
            if obj == 0:
                break
            p += 16 + _pad8(size)
        raise KeyError(f'global heap object {index}')


class H5Dataset:
    def __init__(self, reader, name, messages):
        self._r, self.name = reader, name
        self.shape, self._dtype, self._layout = (), None, None
        for mtype, _, body in messages:
            if mtype == _MSG_DATASPACE:
                self.shape = reader.shape(body)
            elif mtype == _MSG_DATATYPE:
                self._dtype = reader.dtype(body)
            elif mtype == _MSG_FILTERS:
                raise NotImplementedError(f'{name}: filtered (compressed) HDF5 datasets')
            elif mtype == _MSG_LAYOUT:
                version, kind = reader.buf[body], reader.buf[body + 1]
                if version != 3:
                    raise NotImplementedError(f'{name}: HDF5 layout message version {version}')
                if kind == 1:
                    self._layout = ('contiguous', reader.u(body + 2, 8), reader.u(body + 10, 8))
                elif kind == 0:
                    self._layout = ('compact', body + 4, reader.u(body + 2, 2))
                else:
                    raise NotImplementedError(f'{name}: chunked HDF5 datasets')
        if self._dtype is None or self._layout is None:
            raise ValueError(f'{name}: dataset without datatype / layout')

    @property
    def dtype(self):
        return object if isinstance(self._dtype, tuple) else self._dtype

    def read(self):
        """The whole dataset: numpy array, or a list of str for variable-length strings."""
        count = int(np.prod(self.shape)) if self.shape else 1
        _, addr, nbytes = self._layout
        if self._layout[0] == 'contiguous':
            addr = None if addr == _UNDEF else addr + self._r.base
        raw = b'' if addr is None else self._r.buf[addr:addr + nbytes]
        if isinstance(self._dtype, tuple):
            return self._r.vlen_strings(raw, count if addr is not None else 0)
        if addr is None:
            return np.zeros(self.shape, dtype=self._dtype)
        return np.frombuffer(raw, dtype=self._dtype, count=count).reshape(self.shape).copy()

    def __getitem__(self, item):
        return self.read()[item]


class H5Group:
    def __init__(self, reader, name, messages):
        self._r, self.name, self._links, self._attr_msgs = reader, name, {}, []
        for mtype, _, body in messages:
            if mtype == _MSG_SYMBOL_TABLE:
                self._links = reader.group_links(reader.u(body, 8), reader.u(body + 8, 8))
            elif mtype == _MSG_ATTRIBUTE:
                self._attr_msgs.append(body)

    def keys(self):
        return sorted(self._links)

    def __contains__(self, key):
        return key in self._links

    def __len__(self):
        return len(self._links)

    def __getitem__(self, key):
        obj = self
        for part in key.strip('/').split('/'):
            addr = obj._links[part]
            msgs = obj._r.messages(addr)
            path = f'{obj.name}/{part}'
            if any(m[0] == _MSG_SYMBOL_TABLE for m in msgs):
                obj = H5Group(obj._r, path, msgs)
            else:
                obj = H5Dataset(obj._r, path, msgs)
        return obj

    @property
    def attrs(self):
        """Numeric (scalar or 1-D) attributes as a dict."""
        out, r = {}, self._r
        for body in self._attr_msgs:
            if r.buf[body] != 1:
                continue          # other attribute message versions: not written by h5py here
            name_size, type_size, space_size = (r.u(body + 2, 2), r.u(body + 4, 2),
                                                r.u(body + 6, 2))
            p = body + 8
            name = r.buf[p:p + name_size].split(b'\0')[0].decode('utf-8')
            p += _pad8(name_size)
            try:
                dtype = r.dtype(p)
            except NotImplementedError:
                continue
            shape = r.shape(p + _pad8(type_size))
            p += _pad8(type_size) + _pad8(space_size)
            if isinstance(dtype, tuple):
                continue
            count = int(np.prod(shape)) if shape else 1
            value = np.frombuffer(r.buf[p:p + count * dtype.itemsize], dtype=dtype, count=count)
            out[name] = value.reshape(shape).copy() if shape else value[0]
        return out


class H5File(H5Group):
    def __init__(self, path):
        with open(path, 'rb') as fh:
            reader = _Reader(fh.read())
        root = reader.u(reader.root_entry + reader.O, 8)
        super().__init__(reader, '', reader.messages(root))
        self.path = path

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
