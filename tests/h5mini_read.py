"""An independent, minimal reader of the HDF5 structures the host executable writes (version-0 superblock,
old-style groups, contiguous float64 datasets), written from the HDF5 File Format Specification. Test
infrastructure: this image has neither libhdf5 nor h5py, so this reader is the only check of h5mini.hpp."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class File:
    def __init__(self, path):
        self.b = open(path, 'rb').read()
        b = self.b
        if b[:8] != b'\x89HDF\r\n\x1a\n':
            raise H5Error('bad signature')
        sb_ver, fs_ver, rg_ver, _, sh_ver, so, sl, _ = struct.unpack_from('<8B', b, 8)
        if (sb_ver, fs_ver, rg_ver, sh_ver, so, sl) != (0, 0, 0, 0, 8, 8):
            raise H5Error('unsupported superblock')
        self.leaf_k, self.int_k = struct.unpack_from('<HH', b, 16)
        base, fs, eof, drv = struct.unpack_from('<4Q', b, 24)
        if base != 0 or eof != len(b) or fs != UNDEF or drv != UNDEF:
            raise H5Error('bad superblock addresses (eof %d vs file %d)' % (eof, len(b)))
        name_off, hdr, cache, _ = struct.unpack_from('<QQII', b, 56)
        self.root_btree, self.root_heap = struct.unpack_from('<QQ', b, 80)
        if cache != 1 or name_off != 0:
            raise H5Error('root entry is not a cached group')
        self.root = hdr
        msgs = self._messages(hdr)
        st = [m for m in msgs if m[0] == 0x11]
        if len(st) != 1 or struct.unpack('<QQ', st[0][1][:16]) != (self.root_btree, self.root_heap):
            raise H5Error('root symbol table message does not match the superblock scratch space')

    def _messages(self, at):
        b = self.b
        ver, _, n, ref, size = struct.unpack_from('<BBHII', b, at)
        if ver != 1 or at % 8:
            raise H5Error('bad object header at %d' % at)
        pos, end, out = at + 16, at + 16 + size, []
        for _ in range(n):
            typ, sz, flags = struct.unpack_from('<HHB', b, pos)
            if sz % 8 or pos + 8 + sz > end:
                raise H5Error('bad message size')
            out.append((typ, b[pos + 8:pos + 8 + sz], flags))
            pos += 8 + sz
        if pos != end:
            raise H5Error('messages do not fill the header chunk')
        return out

    def _heap_string(self, heap_at, off):
        b = self.b
        if b[heap_at:heap_at + 4] != b'HEAP' or b[heap_at + 4] != 0:
            raise H5Error('bad local heap')
        size, free, seg = struct.unpack_from('<QQQ', b, heap_at + 8)
        # walk the free list: offsets inside the segment, terminated by 1
        f, hops = free, 0
        while f != 1:
            if f + 16 > size or hops > 1000:
                raise H5Error('bad heap free list')
            f, fsize = struct.unpack_from('<QQ', b, seg + f)
            hops += 1
            if fsize < 16:
                raise H5Error('free block too small')
        end = b.index(b'\0', seg + off)
        if end >= seg + size:
            raise H5Error('name runs out of the heap')
        return b[seg + off:end].decode()

    def links(self, header=None):
        """name -> object header address of an old-style group."""
        b = self.b
        msgs = self._messages(self.root if header is None else header)
        st = [m for m in msgs if m[0] == 0x11]
        if not st:
            raise H5Error('not a group')
        btree, heap = struct.unpack('<QQ', st[0][1][:16])
        if b[btree:btree + 4] != b'TREE':
            raise H5Error('bad B-tree node')
        ntype, level, used, left, right = struct.unpack_from('<BBHQQ', b, btree + 4)
        if (ntype, level, left, right) != (0, 0, UNDEF, UNDEF):
            raise H5Error('unsupported B-tree shape')
        out, prev_name = {}, ''
        for e in range(used):
            key_lo, child, key_hi = struct.unpack_from('<QQQ', b, btree + 24 + 16 * e)
            if b[child:child + 4] != b'SNOD' or b[child + 4] != 1:
                raise H5Error('bad symbol table node')
            n = struct.unpack_from('<H', b, child + 6)[0]
            if n > 2 * self.leaf_k:
                raise H5Error('too many symbols in a node')
            names = []
            for k in range(n):
                noff, hdr, cache, _ = struct.unpack_from('<QQII', b, child + 8 + 40 * k)
                names.append(self._heap_string(heap, noff))
                out[names[-1]] = hdr
            if names != sorted(names) or (names and names[0] <= prev_name and e > 0):
                raise H5Error('symbols not sorted')
            if names and self._heap_string(heap, key_hi) != names[-1]:
                raise H5Error('B-tree key does not name the last symbol of its child')
            if self._heap_string(heap, key_lo) >= (names[0] if names else '~'):
                raise H5Error('B-tree left key is not below the child')
            prev_name = names[-1] if names else prev_name
        return out

    def dataset(self, header):
        msgs = {m[0]: m[1] for m in self._messages(header)}
        for need in (0x1, 0x3, 0x8):
            if need not in msgs:
                raise H5Error('dataset message %#x missing' % need)
        sp = msgs[0x1]
        if sp[0] != 1 or sp[2] != 0:
            raise H5Error('unsupported dataspace')
        dims = struct.unpack_from('<%dQ' % sp[1], sp, 8)
        dt = msgs[0x3]
        if bytes(dt[:20]) != bytes([0x11, 0x20, 0x3f, 0, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0]):
            raise H5Error('datatype is not IEEE F64LE')
        lay = msgs[0x8]
        if lay[0] != 3 or lay[1] != 1:
            raise H5Error('layout is not contiguous v3')
        addr, size = struct.unpack_from('<QQ', lay, 2)
        n = int(np.prod(dims)) if dims else 1
        if size != 8 * n or addr % 8 or addr + size > len(self.b):
            raise H5Error('bad layout extent')
        return np.frombuffer(self.b, dtype='<f8', count=n, offset=addr).reshape(dims).copy()

    def read(self, path):
        """read('/part') or read('/grp/part')"""
        parts = [p for p in path.split('/') if p]
        hdr = None
        for p in parts[:-1]:
            hdr = self.links(hdr)[p]
        return self.dataset(self.links(hdr)[parts[-1]])
