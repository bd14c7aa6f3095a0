"""ROOT files without ROOT: a writer and a reader for FLAT TTrees (one fixed-size number per branch and entry),
the wire format either side of the photon path.

The reference writes the rows of a PhaseSpaceActor as a TTree through Geant4's analysis manager
(`GateDigiCollectionsRootManager.cpp`, column types `GateDigiCollection.cpp:240-271`: D = double, I = int,
3 = three doubles `name_X/_Y/_Z`) and reads phase-space files back through `uproot` (`sources/phspsources.py:59-75`:
first tree of the file, branches by name).  Neither ROOT nor uproot exists in every deployment of this engine, so
this module states the container itself, from the published description of the format (ROOT `TFile`, `TKey`,
`TDirectory`, `TTree` v20, `TBranch` v13, `TLeaf*`, `TBasket` v3 streamers; big-endian):

    [file header, 100 bytes][top directory record][basket records ...][TTree record][StreamerInfo record]
    [keys list][free-segments record]

* `write_tree(path, name, columns)`: float64 / float32 / int32 / int64 / int16 / int8 / bool / string columns -> one
  branch and one leaf each (strings: `TLeafC` with an entry-offset table per basket), data in basket records of at most `BASKET_ENTRIES` entries, optionally zlib-compressed ("ZL"
  blocks).  Files beyond 2 GB switch to the 64-bit seek layout (header version + 1000000, key version + 1000).
  Class versions are the ones ROOT 6.14 ... 6.3x compile in (a StreamerInfo list is present but empty).
* `read_tree(path, name=None, keys=None)`: the same subset back -- any uncompressed, zlib- or lzma-compressed flat
  tree with fixed-size leaves (TTree v19 / v20, TBranch v12 / v13), whoever wrote it.

What is verified here: every file the writer produces is read back bit for bit by the independent reader
(`tests/test_rootio.py`, including compression, several baskets and the 64-bit layout), and a PhaseSpaceActor ->
`.root` -> PhaseSpaceSource chain reproduces the `.npz` chain.  No ROOT installation was available to open the
files with: byte-level agreement with ROOT's own reader is by construction from the format description, not by test.
"""
import struct
import time
import uuid
import zlib

import numpy as np

K_BYTECOUNT = 0x40000000
K_NEW_CLASS = 0xFFFFFFFF
K_CLASS_MASK = 0x80000000
K_MAP_OFFSET = 2
BASKET_ENTRIES = 1 << 20
ROOT_VERSION = 62400          # "written by ROOT 6.24/00"
START_BIG = 2000000000        # kStartBigFile: seeks beyond this are 64 bit

# numpy dtype -> (leaf class, leaf title suffix, bytes per entry, unsigned)
LEAVES = {
    "float64": ("TLeafD", "D", 8, False), "float32": ("TLeafF", "F", 4, False),
    "int64": ("TLeafL", "L", 8, False), "uint64": ("TLeafL", "l", 8, True),
    "int32": ("TLeafI", "I", 4, False), "uint32": ("TLeafI", "i", 4, True),
    "int16": ("TLeafS", "S", 2, False), "uint16": ("TLeafS", "s", 2, True),
    "int8": ("TLeafB", "B", 1, False), "uint8": ("TLeafB", "b", 1, True),
    "bool": ("TLeafO", "O", 1, False),
}
LEAF_DTYPE = {"TLeafD": ">f8", "TLeafF": ">f4", "TLeafL": ">i8", "TLeafI": ">i4", "TLeafS": ">i2", "TLeafB": ">i1", "TLeafO": "?"}
LEAF_MINMAX = {"TLeafD": ">dd", "TLeafF": ">ff", "TLeafL": ">qq", "TLeafI": ">ii", "TLeafS": ">hh", "TLeafB": ">bb", "TLeafO": ">??"}


def _tstring(s):
    b = s.encode("utf-8")
    return (bytes([len(b)]) if len(b) < 255 else b"\xff" + struct.pack(">i", len(b))) + b


def _datime(t=None):
    """TDatime: (year - 1995) << 26 | month << 22 | day << 17 | hour << 12 | minute << 6 | second"""
    tm = time.localtime(t)
    return ((tm.tm_year - 1995) << 26) | (tm.tm_mon << 22) | (tm.tm_mday << 17) | (tm.tm_hour << 12) | (tm.tm_min << 6) | min(tm.tm_sec, 59)


def _key_header(nbytes, objlen, keylen, seek_key, seek_pdir, classname, name, title, datime, cycle=1, big=False):
    """TKey::FillBuffer; `big`: 64-bit seeks, version + 1000"""
    h = struct.pack(">ihiIhh", nbytes, 1004 if big else 4, objlen, datime, keylen, cycle)
    h += struct.pack(">qq", seek_key, seek_pdir) if big else struct.pack(">ii", seek_key, seek_pdir)
    return h + _tstring(classname) + _tstring(name) + _tstring(title)


def _key_length(classname, name, title, big=False):
    return 18 + (16 if big else 8) + len(_tstring(classname)) + len(_tstring(name)) + len(_tstring(title))


def _compress(data, level):
    """ROOT's framing of zlib streams: blocks of < 16 MB, each 'ZL' 0x08 + 3-byte compressed and uncompressed sizes
    (little endian).  A block that does not shrink makes the whole buffer be stored raw, as ROOT does."""
    out = []
    for o in range(0, len(data), 0xFFFFFF):
        chunk = data[o:o + 0xFFFFFF]
        z = zlib.compress(chunk, level)
        if len(z) + 9 >= len(chunk):
            return data
        c, u = len(z), len(chunk)
        out.append(b"ZL\x08" + bytes([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF, u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF]) + z)
    return b"".join(out)


def _decompress(data, objlen):
    if len(data) == objlen:
        return data
    out, o = [], 0
    while o < len(data):
        magic = data[o:o + 2]
        c = data[o + 3] | (data[o + 4] << 8) | (data[o + 5] << 16)
        u = data[o + 6] | (data[o + 7] << 8) | (data[o + 8] << 16)
        blk = data[o + 9:o + 9 + c]
        if magic == b"ZL":
            out.append(zlib.decompress(blk))
        elif magic == b"XZ":
            import lzma
            out.append(lzma.decompress(blk))
        else:
            raise ValueError(f"ROOT compression '{magic.decode('latin1')}' (LZ4 / ZSTD / old zlib) is not supported by this reader")
        if len(out[-1]) != u:
            raise ValueError("corrupt compressed block")
        o += 9 + c
    return b"".join(out)


class _Buf:
    """serialisation buffer of one keyed object: positions count from the start of the key (TBufferFile of a TKey),
    which is what class and object references inside the object refer to"""

    def __init__(self, keylen):
        self.b = bytearray()
        self.keylen = keylen
        self.classes = {}

    @property
    def pos(self):
        return self.keylen + len(self.b)

    def add(self, fmt, *v):
        self.b += struct.pack(">" + fmt, *v)

    def raw(self, b):
        self.b += b

    def begin(self, version=None):
        """byte count (filled by end) + class version"""
        p = len(self.b)
        self.b += b"\0\0\0\0"
        if version is not None:
            self.add("h", version)
        return p

    def end(self, p):
        struct.pack_into(">I", self.b, p, (len(self.b) - p - 4) | K_BYTECOUNT)

    def tobject(self):
        self.add("hII", 1, 0, 0x03000000)

    def tnamed(self, name, title):
        p = self.begin(1)
        self.tobject()
        self.raw(_tstring(name) + _tstring(title))
        self.end(p)

    def object_any(self, classname, payload):
        """TBufferFile::WriteObjectAny: byte count, class tag (new class: -1 + name; known class: reference), object.
        Returns the reference later pointers to this object use."""
        p = self.begin()
        if classname in self.classes:
            self.add("I", self.classes[classname] | K_CLASS_MASK)
        else:
            self.classes[classname] = self.pos + K_MAP_OFFSET
            self.add("I", K_NEW_CLASS)
            self.raw(classname.encode() + b"\0")
        payload(self)
        self.end(p)
        return self.keylen + p + K_MAP_OFFSET

    def objarray(self, writers):
        """TObjArray of objects written by `writers` (callables) or of references (ints)"""
        p = self.begin(3)
        self.tobject()
        self.raw(_tstring(""))
        self.add("ii", len(writers), 0)
        for w in writers:
            if isinstance(w, int):
                self.add("I", w)
            else:
                w(self)
        self.end(p)

    def iofeatures(self):
        # ROOT::TIOFeatures, class version 0 -> streamed with its checksum; no feature bits
        self.raw(b"\x40\x00\x00\x07\x00\x00\x1a\xa1\x2f\x10\x00")


def write_tree(path, tree_name, columns, title="", compression=0, basket_entries=BASKET_ENTRIES, big=None):
    """Write dict name -> 1-D numpy array (all of one length) as a flat TTree `tree_name` into a new ROOT file.
    compression: 0 = stored, 1..9 = zlib level (file setting 100 + level).  big: force the 64-bit layout (tests)."""
    cols = {}
    n = None
    for k, v in columns.items():
        a = np.asarray(v)
        if a.ndim != 1:
            raise ValueError(f"branch '{k}': only one number per entry (got shape {a.shape})")
        if a.dtype.kind in "US":
            a = np.char.encode(a, "utf-8") if a.dtype.kind == "U" else a     # -> bytes: a TLeafC branch
        elif a.dtype.name not in LEAVES:
            raise ValueError(f"branch '{k}': dtype {a.dtype} has no ROOT leaf here ({sorted(LEAVES)} or strings)")
        if n is None:
            n = a.shape[0]
        if a.shape[0] != n:
            raise ValueError("all branches of a flat tree have the same number of entries")
        cols[k] = a
    n = int(n or 0)
    fname = str(path).split("/")[-1]
    now = _datime()
    total = 400 + sum(a.nbytes for a in cols.values()) + 4096 * max(len(cols), 1)
    if big is None:
        big = total > START_BIG
    fcompress = 100 + compression if compression else 0

    with open(path, "wb") as f:
        # ---- header and top directory are rewritten at the end (fixed size)
        dir_keylen = _key_length("TFile", fname, "", big)
        nbytes_name = dir_keylen + len(_tstring(fname)) + len(_tstring(""))
        dir_len = 2 + 4 + 4 + 4 + 4 + (24 if big else 12) + 18 + (0 if big else 12)
        f.write(b"\0" * (100 + nbytes_name + dir_len))

        # ---- baskets: one TKey record each, class TBasket, name = branch, title = tree
        branches = []
        for k, a in cols.items():
            strings = a.dtype.kind == "S"
            if strings:
                # TLeafC: every entry is one length byte (255 + int32 for longer strings) and the characters; the
                # basket ends with the table of entry offsets (TBasket::WriteBuffer: count, fKeylen + offset of each
                # entry, one spare word), fLast pointing at that table
                cls, suffix, size, uns = "TLeafC", "C", 1, False
                maxlen = max((len(x) for x in a), default=0)
            else:
                cls, suffix, size, uns = LEAVES[a.dtype.name]
                be = a.astype(LEAF_DTYPE[cls], copy=False) if cls != "TLeafO" else a
            keylen = _key_length("TBasket", k, tree_name, big) + 2 + 4 + 4 + 4 + 4 + 1
            info = {"name": k, "cls": cls, "title": f"{k}/{suffix}", "size": size, "unsigned": uns,
                    "seek": [], "nbytes": [], "first": [], "tot": 0, "zip": 0,
                    "flen": (maxlen + 1) if strings else 1, "offset_len": 1000 if strings else 0}
            for e0 in range(0, n, basket_entries):
                e1 = min(n, e0 + basket_entries)
                tail = b""
                if strings:
                    parts, offs, pos = [], [], 0
                    for x in a[e0:e1]:
                        x = bytes(x)
                        offs.append(pos)
                        enc = (bytes([len(x)]) if len(x) < 255 else b"\xff" + struct.pack(">i", len(x))) + x
                        parts.append(enc)
                        pos += len(enc)
                    data = b"".join(parts)
                else:
                    data = be[e0:e1].tobytes()
                pos = f.tell()
                kbig = big or pos > START_BIG
                kl = keylen + (8 if kbig and not big else 0)
                last = kl + len(data)
                if strings:
                    tail = struct.pack(f">i{e1 - e0 + 1}i", e1 - e0 + 1, *[kl + o for o in offs], 0)
                    data += tail
                z = _compress(data, compression) if compression else data
                hdr = _key_header(kl + len(z), len(data), kl, pos, 100, "TBasket", k, tree_name, now, cycle=0, big=kbig)
                hdr += struct.pack(">hiiiib", 3, 32000, 1000 if strings else size, e1 - e0, last, 0)
                assert len(hdr) == kl
                f.write(hdr)
                f.write(z)
                info["seek"].append(pos); info["nbytes"].append(kl + len(z)); info["first"].append(e0)
                info["tot"] += kl + len(data); info["zip"] += kl + len(z)
            branches.append(info)

        # ---- the TTree object
        seek_tree = f.tell()
        tbig = big or seek_tree > START_BIG
        tkeylen = _key_length("TTree", tree_name, title, tbig)
        B = _Buf(tkeylen)
        p_tree = B.begin(20)
        B.tnamed(tree_name, title)
        p = B.begin(2); B.add("hhh", 602, 1, 1); B.end(p)            # TAttLine
        p = B.begin(2); B.add("hh", 0, 1001); B.end(p)               # TAttFill
        p = B.begin(2); B.add("hhf", 1, 1, 1.0); B.end(p)            # TAttMarker
        tot = sum(b["tot"] for b in branches)
        zp = sum(b["zip"] for b in branches)
        B.add("qqqqq", n, tot, zp, 0, 0)                             # fEntries fTotBytes fZipBytes fSavedBytes fFlushedBytes
        B.add("diiii", 1.0, 0, 25, 0, 1000)                          # fWeight fTimerInterval fScanField fUpdate fDefaultEntryOffsetLen
        B.add("i", 0)                                                # fNClusterRange
        B.add("qqqqqq", 1000000000000, 1000000000000, 0, -300000000, -30000000, 1000000)
        B.raw(b"\0\0")                                               # fClusterRangeEnd, fClusterSize: no arrays
        B.iofeatures()
        leaf_refs = []

        def branch_writer(info):
            def leaf_payload(b):
                p1 = b.begin(1)                                       # TLeafX
                p2 = b.begin(2)                                       # TLeaf
                b.tnamed(info["name"], info["name"])
                b.add("iii??I", info["flen"], info["size"], 0, False, info["unsigned"], 0)   # fLen fLenType fOffset fIsRange fIsUnsigned fLeafCount
                b.end(p2)
                if info["cls"] == "TLeafC":
                    b.add("ii", 0, info["flen"])                      # fMinimum, fMaximum (longest string + 1)
                else:
                    b.raw(struct.pack(LEAF_MINMAX[info["cls"]], 0, 0))
                b.end(p1)

            def payload(b):
                nb = len(info["seek"])
                maxb = max(10, nb + 1)
                p1 = b.begin(13)                                      # TBranch
                b.tnamed(info["name"], info["title"])
                p2 = b.begin(2); b.add("hh", 0, 1001); b.end(p2)      # TAttFill
                b.add("iiiiq", fcompress, 32000, info["offset_len"], nb, n)   # fCompress fBasketSize fEntryOffsetLen fWriteBasket fEntryNumber
                b.iofeatures()
                b.add("iiiqqqq", 0, maxb, 0, n, 0, info["tot"], info["zip"])   # fOffset fMaxBaskets fSplitLevel fEntries fFirstEntry fTotBytes fZipBytes
                b.objarray([])                                        # fBranches
                b.objarray([lambda bb: leaf_refs.append(bb.object_any(info["cls"], leaf_payload))])   # fLeaves
                b.objarray([])                                        # fBaskets: all written out
                b.raw(b"\x01" + np.asarray(info["nbytes"] + [0] * (maxb - nb), ">i4").tobytes())      # fBasketBytes
                b.raw(b"\x01" + np.asarray(info["first"] + [n] + [0] * (maxb - nb - 1), ">i8").tobytes())   # fBasketEntry
                b.raw(b"\x01" + np.asarray(info["seek"] + [0] * (maxb - nb), ">i8").tobytes())        # fBasketSeek
                b.raw(_tstring(""))                                   # fFileName
                b.end(p1)
            return lambda b: b.object_any("TBranch", payload)

        B.objarray([branch_writer(i) for i in branches])             # fBranches
        B.objarray(leaf_refs)                                         # fLeaves: the same leaf objects, by reference
        B.add("I", 0)                                                # fAliases
        B.add("i", 0); B.add("i", 0)                                 # fIndexValues (TArrayD), fIndex (TArrayI)
        B.add("IIII", 0, 0, 0, 0)                                    # fTreeIndex fFriends fUserInfo fBranchRef
        B.end(p_tree)
        obj = bytes(B.b)
        zobj = _compress(obj, compression) if compression else obj
        tree_key = _key_header(tkeylen + len(zobj), len(obj), tkeylen, seek_tree, 100, "TTree", tree_name, title, now, 1, tbig)
        f.write(tree_key + zobj)

        # ---- StreamerInfo record: an empty TList (the class versions written are the ones ROOT compiles in)
        seek_info = f.tell()
        ibig = big or seek_info > START_BIG
        ikeylen = _key_length("TList", "StreamerInfo", "Doubly linked list", ibig)
        L = _Buf(ikeylen)
        p = L.begin(5); L.tobject(); L.raw(_tstring("")); L.add("i", 0); L.end(p)
        info_rec = _key_header(ikeylen + len(L.b), len(L.b), ikeylen, seek_info, 100, "TList", "StreamerInfo", "Doubly linked list", now, 1, ibig) + bytes(L.b)
        f.write(info_rec)

        # ---- keys list of the top directory
        seek_keys = f.tell()
        kbig = big or seek_keys > START_BIG
        klen = _key_length("TFile", fname, "", kbig)
        body = struct.pack(">i", 1) + tree_key
        f.write(_key_header(klen + len(body), len(body), klen, seek_keys, 100, "TFile", fname, "", now, 1, kbig) + body)
        nbytes_keys = klen + len(body)

        # ---- free segments: one, from the end of the file
        seek_free = f.tell()
        fbig = big or seek_free > START_BIG
        flen = _key_length("TFile", fname, "", fbig)
        nfree_bytes = flen + (2 + 16 if fbig else 2 + 8)
        end = seek_free + nfree_bytes
        if fbig:
            body = struct.pack(">hqq", 1001, end, START_BIG if end < START_BIG else end + 1000000000)
        else:
            body = struct.pack(">hii", 1, end, START_BIG)
        f.write(_key_header(nfree_bytes, len(body), flen, seek_free, 100, "TFile", fname, "", now, 1, fbig) + body)
        assert f.tell() == end

        # ---- header + top directory
        f.seek(0)
        uid = uuid.uuid1().bytes
        hbig = big or end > START_BIG
        if hbig:
            hdr = struct.pack(">4siiqqiiiBiqi", b"root", ROOT_VERSION + 1000000, 100, end, seek_free, nfree_bytes, 1, nbytes_name, 8,
                              fcompress, seek_info, len(info_rec))
        else:
            hdr = struct.pack(">4siiiiiiiBiii", b"root", ROOT_VERSION, 100, end, seek_free, nfree_bytes, 1, nbytes_name, 4,
                              fcompress, seek_info, len(info_rec))
        hdr += struct.pack(">h", 1) + uid
        f.write(hdr + b"\0" * (100 - len(hdr)))
        d = _key_header(nbytes_name + dir_len, dir_len + len(_tstring(fname)) + len(_tstring("")), dir_keylen, 100, 0, "TFile", fname, "", now, 1, big)
        d += _tstring(fname) + _tstring("")
        if big:
            d += struct.pack(">hIIiiqqq", 1005, now, now, nbytes_keys, nbytes_name, 100, 0, seek_keys)
        else:
            d += struct.pack(">hIIiiiii", 5, now, now, nbytes_keys, nbytes_name, 100, 0, seek_keys)
        d += struct.pack(">h", 1) + uid
        if not big:
            d += b"\0" * 12
        assert len(d) == nbytes_name + dir_len, (len(d), nbytes_name, dir_len)
        f.write(d)
    return n


# ------------------------------------------------------------------------------------------------ reader

class _Cur:
    def __init__(self, b, pos=0, origin=0):
        self.b, self.p, self.origin = b, pos, origin   # origin: value of `pos` references count from

    def get(self, fmt):
        v = struct.unpack_from(">" + fmt, self.b, self.p)
        self.p += struct.calcsize(">" + fmt)
        return v if len(v) > 1 else v[0]

    def tstring(self):
        n = self.b[self.p]
        self.p += 1
        if n == 255:
            n = self.get("i")
        s = self.b[self.p:self.p + n]
        self.p += n
        return bytes(s).decode("utf-8", "replace")

    def cstring(self):
        e = self.b.index(b"\0", self.p)
        s = bytes(self.b[self.p:e]).decode()
        self.p = e + 1
        return s

    def version(self):
        """-> (end position of the object or None, class version)"""
        end = None
        c = struct.unpack_from(">I", self.b, self.p)[0]
        if c & K_BYTECOUNT:
            self.p += 4
            end = self.p + (c & ~K_BYTECOUNT)
        return end, self.get("h")

    def skip_versioned(self):
        end, _ = self.version()
        if end is None:
            raise ValueError("object without a byte count")
        self.p = end


def _read_key(c):
    nbytes, ver, objlen, datime, keylen, cycle = c.get("ihiIhh")
    seek_key, seek_pdir = c.get("qq") if ver > 1000 else c.get("ii")
    return {"nbytes": nbytes, "objlen": objlen, "keylen": keylen, "cycle": cycle, "seek": seek_key,
            "class": c.tstring(), "name": c.tstring(), "title": c.tstring()}


def _read_tnamed(c):
    end, _ = c.version()
    c.get("hII")
    name, title = c.tstring(), c.tstring()
    if end is not None:
        c.p = end
    return name, title


def _read_iofeatures(c):
    c.skip_versioned()


def _read_objarray(c, classes, reader):
    """TObjArray -> list of reader(c, classname) results (None for null pointers / references)"""
    end, _ = c.version()
    c.get("hII")
    c.tstring()
    n, _low = c.get("ii")
    out = []
    for _ in range(n):
        start = c.p
        tag = c.get("I")
        if tag == 0:
            out.append(None)
            continue
        if not (tag & K_BYTECOUNT):
            out.append(("ref", tag))   # reference to an object streamed earlier
            continue
        oend = c.p + (tag & ~K_BYTECOUNT)
        ctag_pos = c.p
        ctag = c.get("I")
        if ctag == K_NEW_CLASS:
            cname = c.cstring()
            classes[ctag_pos - c.origin + K_MAP_OFFSET] = cname
        else:
            cname = classes[ctag & ~K_CLASS_MASK]
        out.append(reader(c, cname, start - c.origin + K_MAP_OFFSET))
        c.p = oend
    if end is not None:
        c.p = end
    return out


def _read_leaf(c, cname, ref):
    end, _ = c.version()          # TLeafX
    lend, _ = c.version()         # TLeaf
    name, title = _read_tnamed(c)
    flen, flentype, _off, _isrange, unsigned = c.get("iii??")
    count = c.get("I")
    if lend is not None:
        c.p = lend
    return {"class": cname, "name": name, "len": flen, "size": flentype, "unsigned": unsigned, "count": count, "ref": ref}


def _read_branch(c, cname, ref, classes):
    if cname != "TBranch":
        raise ValueError(f"branch class {cname}: only plain TBranch (flat trees) is supported")
    end, ver = c.version()
    if ver not in (12, 13):
        raise ValueError(f"TBranch version {ver} is not supported (12, 13)")
    name, title = _read_tnamed(c)
    c.skip_versioned()                                   # TAttFill
    fcompress, _bsize, entry_offset_len, write_basket, _entry_number = c.get("iiiiq")
    if ver >= 13:
        _read_iofeatures(c)
    _offset, max_baskets, _split, entries, _first, _tot, _zip = c.get("iiiqqqq")
    sub = _read_objarray(c, classes, lambda cc, cn, r: _read_branch(cc, cn, r, classes))
    leaves = _read_objarray(c, classes, _read_leaf)
    baskets = _read_objarray(c, classes, lambda cc, cn, r: cn)
    if any(b is not None for b in baskets):
        raise ValueError(f"branch '{name}' keeps a basket inside the tree record (file not flushed): not supported")
    arrays = []
    for fmt in (">i4", ">i8", ">i8"):
        flag = c.get("b")
        a = np.frombuffer(c.b, fmt, max_baskets, c.p) if flag else np.zeros(0, fmt)
        c.p += a.nbytes
        arrays.append(a)
    c.tstring()
    if sub:
        raise ValueError(f"branch '{name}' has sub-branches: not a flat tree")
    if len(leaves) != 1 or not isinstance(leaves[0], dict) or leaves[0]["count"] != 0 or \
            (leaves[0]["class"] not in LEAF_DTYPE and leaves[0]["class"] != "TLeafC"):
        raise ValueError(f"branch '{name}': only one fixed-size leaf per branch is supported")
    return {"name": name, "title": title, "leaf": leaves[0], "n_baskets": write_basket, "entries": entries,
            "basket_bytes": arrays[0], "basket_entry": arrays[1], "basket_seek": arrays[2]}


def _read_strings(f, b, entries):
    """a TLeafC branch: the entries of a basket lie between the offsets of its table (after fLast), each one a length
    byte (255 + int32 for long strings) and the characters"""
    vals = []
    for i in range(b["n_baskets"]):
        f.seek(int(b["basket_seek"][i]))
        raw = f.read(int(b["basket_bytes"][i]))
        bc = _Cur(raw)
        bk = _read_key(bc)
        _bver, _bufsize, _nevsize, nev, last = bc.get("hiiii")
        data = _decompress(raw[bk["keylen"]:], bk["objlen"])
        border = last - bk["keylen"]
        table = np.frombuffer(data, ">i4", offset=border)
        offs = (table[1:1 + nev].astype(np.int64) - bk["keylen"]).tolist() + [border]
        for j in range(nev):
            s = data[offs[j]:offs[j + 1]]
            k = 1
            if s[:1] == b"\xff":
                k = 5
            vals.append(bytes(s[k:]).decode("utf-8", "replace"))
    if len(vals) != entries:
        raise ValueError(f"branch '{b['name']}': {len(vals)} entries read, tree has {entries}")
    return np.array(vals, dtype=str) if vals else np.zeros(0, dtype="U1")


def list_trees(path):
    """names of the TTrees in the top directory of the file, in key order"""
    with open(path, "rb") as f:
        return [k["name"] for k in _keys(f) if k["class"] == "TTree"]


def _keys(f):
    head = f.read(100)
    if head[:4] != b"root":
        raise ValueError("not a ROOT file")
    c = _Cur(head, 4)
    version, begin = c.get("ii")
    if version >= 1000000:
        _end, _seek_free, _nbf, _nfree, nbytes_name = c.get("qqiii")
    else:
        _end, _seek_free, _nbf, _nfree, nbytes_name = c.get("iiiii")
    f.seek(begin + nbytes_name)
    d = _Cur(f.read(64))
    dver = d.get("h")
    d.get("II")
    _nbytes_keys, _nbn = d.get("ii")
    _seek_dir, _seek_parent, seek_keys = d.get("qqq") if dver > 1000 else d.get("iii")
    f.seek(seek_keys)
    kh = f.read(4096)
    c = _Cur(kh)
    top = _read_key(c)
    f.seek(seek_keys)
    c = _Cur(f.read(top["nbytes"]), top["keylen"])
    return [_read_key(c) for _ in range(c.get("i"))]


def read_tree(path, name=None, keys=None):
    """-> (dict branch name -> numpy array (native byte order), list of all branch names).  name: tree to read
    (default: the first TTree key, as the reference's reader takes the first key); keys: branches wanted (those present)."""
    with open(path, "rb") as f:
        trees = [k for k in _keys(f) if k["class"] == "TTree" and (name is None or k["name"] == name)]
        if not trees:
            raise KeyError(f"no TTree{'' if name is None else ' ' + repr(name)} in {path}")
        first = trees[0]["name"]
        k = max((t for t in trees if t["name"] == first), key=lambda t: t["cycle"])
        f.seek(k["seek"])
        rec = f.read(k["nbytes"])
        obj = _decompress(rec[k["keylen"]:], k["objlen"])
        # references inside the object count from the start of the key
        c = _Cur(bytes(k["keylen"]) + obj, k["keylen"], 0)
        _end, ver = c.version()
        if ver not in (19, 20):
            raise ValueError(f"TTree version {ver} is not supported (19, 20)")
        _read_tnamed(c)
        for _ in range(3):
            c.skip_versioned()                               # TAttLine, TAttFill, TAttMarker
        entries = c.get("qqqqq")[0]
        c.get("diiii")
        ncluster = c.get("i")
        c.get("qqqqqq")
        for _ in range(2):                                   # fClusterRangeEnd, fClusterSize
            if c.get("b"):
                c.p += 8 * ncluster
        if ver >= 20:
            _read_iofeatures(c)
        classes = {}
        branches = _read_objarray(c, classes, lambda cc, cn, r: _read_branch(cc, cn, r, classes))
        avail = [b["name"] for b in branches]
        out = {}
        for b in branches:
            if keys is not None and b["name"] not in keys:
                continue
            leaf = b["leaf"]
            if leaf["class"] == "TLeafC":
                out[b["name"]] = _read_strings(f, b, entries)
                continue
            if leaf["class"] not in LEAF_DTYPE:
                raise ValueError(f"branch '{b['name']}': leaf class {leaf['class']} is not supported")
            dt = np.dtype(LEAF_DTYPE[leaf["class"]])
            if leaf["unsigned"] and dt.kind == "i":
                dt = np.dtype(">u" + str(dt.itemsize))
            parts = []
            for i in range(b["n_baskets"]):
                f.seek(int(b["basket_seek"][i]))
                raw = f.read(int(b["basket_bytes"][i]))
                bc = _Cur(raw)
                bk = _read_key(bc)
                _bver, _bufsize, _nevsize, nev, last = bc.get("hiiii")
                data = _decompress(raw[bk["keylen"]:], bk["objlen"])
                parts.append(np.frombuffer(data, dt, nev * leaf["len"]))
            a = np.concatenate(parts) if parts else np.zeros(0, dt)
            a = a.astype(dt.newbyteorder("="))
            if leaf["len"] > 1:
                a = a.reshape(-1, leaf["len"])
            if a.shape[0] != entries:
                raise ValueError(f"branch '{b['name']}': {a.shape[0]} entries read, tree has {entries}")
            out[b["name"]] = a
        return out, avail
