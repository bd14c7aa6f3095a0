"""The dependency-free ROOT TTree writer / reader (opengate_b200/rootio.py): every file the writer makes is read back
bit for bit by the independent reader; structural known answers of the container (header, key, directory, basket
records) are checked against the published layout."""
import struct

import numpy as np
import pytest

from opengate_b200 import rootio


def columns(n, seed=0):
    r = np.random.default_rng(seed)
    return {
        "KineticEnergy": r.random(n), "Weight": r.random(n).astype(np.float32), "EventID": r.integers(0, 2**31 - 1, n).astype(np.int32),
        "PrePosition_X": r.normal(size=n), "PrePosition_Y": r.normal(size=n), "PrePosition_Z": r.normal(size=n),
        "big": r.integers(-2**62, 2**62, n), "PDGCode": np.full(n, 22, np.int32), "flag": r.integers(0, 2, n).astype(bool),
        "s16": r.integers(-30000, 30000, n).astype(np.int16), "u8": r.integers(0, 255, n).astype(np.uint8),
        "ParticleName": np.array(["gamma", "e-", "", "x" * 300])[r.integers(0, 4, n)] if n else np.zeros(0, dtype="U5"),
    }


@pytest.mark.parametrize("n,compression,basket,big", [(0, 0, 1000, False), (1, 0, 1000, False), (1234, 0, 500, False),
                                                      (5000, 6, 1024, False), (777, 0, 100, True), (3000, 1, 4096, True)])
def test_round_trip(tmp_path, n, compression, basket, big):
    cols = columns(n)
    path = tmp_path / "t.root"
    assert rootio.write_tree(path, "PhaseSpace", cols, title="phsp", compression=compression, basket_entries=basket, big=big) == n
    assert rootio.list_trees(path) == ["PhaseSpace"]
    out, avail = rootio.read_tree(path)
    assert avail == list(cols)
    for k, v in cols.items():
        assert (out[k].dtype == v.dtype or v.dtype.kind == "U") and np.array_equal(out[k], v), k
    sub, _ = rootio.read_tree(path, "PhaseSpace", keys=["Weight", "big", "nope"])
    assert list(sub) == ["Weight", "big"]
    with pytest.raises(KeyError):
        rootio.read_tree(path, "other")


def test_container_layout_known_answers(tmp_path):
    path = tmp_path / "k.root"
    x = np.arange(10, dtype=np.float64)
    rootio.write_tree(path, "T", {"x": x})
    b = path.read_bytes()
    # file header (TFile::WriteHeader): magic, version, fBEGIN = 100, fEND = file size, fUnits = 4 for 32-bit seeks
    magic, version, begin, end, seek_free, nbytes_free, nfree, nbytes_name, units = struct.unpack_from(">4siiiiiiiB", b, 0)
    assert magic == b"root" and version == rootio.ROOT_VERSION and begin == 100 and end == len(b) and units == 4 and nfree == 1
    assert seek_free + nbytes_free == len(b)
    # the top directory record starts at fBEGIN with a TKey of class TFile named after the file
    nbytes, kver, objlen, _dt, keylen, cycle, seek_key, seek_pdir = struct.unpack_from(">ihiIhhii", b, 100)
    assert kver == 4 and seek_key == 100 and seek_pdir == 0 and nbytes == keylen + objlen
    assert b[126:127 + 5] == b"\x05TFile" and b[132:133 + 6] == b"\x06k.root"
    # TDirectory part: 60 bytes (TDirectoryFile::Sizeof), fSeekDir = 100, fSeekKeys points at a TFile key
    dver, _c, _m, nbytes_keys, nbn, seek_dir, seek_parent, seek_keys = struct.unpack_from(">hIIiiiii", b, 100 + nbytes_name)
    assert dver == 5 and nbn == nbytes_name and seek_dir == 100 and seek_parent == 0 and nbytes == nbytes_name + 60
    assert struct.unpack_from(">i", b, seek_keys)[0] == nbytes_keys
    # first basket record: right after the directory; class TBasket, name = branch, title = tree; big-endian payload
    o = 100 + nbytes
    bn, bver, bobjlen, _dt, bkeylen, _cyc, bseek, bpdir = struct.unpack_from(">ihiIhhii", b, o)
    assert bseek == o and bpdir == 100 and bobjlen == 80 and bn == bkeylen + 80
    assert b[o + 26:o + 26 + 8] == b"\x07TBasket" and b[o + 34:o + 36] == b"\x01x" and b[o + 36:o + 38] == b"\x01T"
    tver, bufsize, nevsize, nev, last, flag = struct.unpack_from(">hiiiib", b, o + 38)
    assert (tver, nevsize, nev, last, flag) == (3, 8, 10, bkeylen + 80, 0) and o + 38 + 19 == o + bkeylen
    assert np.array_equal(np.frombuffer(b, ">f8", 10, o + bkeylen), x)
    # the TTree object: byte count | kByteCountMask, version 20, TNamed with the tree's name
    t = o + bn
    tn, _v, tobjlen, _dt, tkeylen = struct.unpack_from(">ihiIh", b, t)
    cnt, tv = struct.unpack_from(">Ih", b, t + tkeylen)
    assert cnt & 0x40000000 and (cnt & 0x3FFFFFFF) == tobjlen - 4 and tv == 20
    assert b"TBranch\x00" in b[t:t + tn] and b"TLeafD\x00" in b[t:t + tn]


def test_compressed_blocks_have_roots_framing(tmp_path):
    path = tmp_path / "z.root"
    rootio.write_tree(path, "T", {"x": np.zeros(100000)}, compression=4)
    b = path.read_bytes()
    i = b.index(b"ZL\x08")
    c = b[i + 3] | (b[i + 4] << 8) | (b[i + 5] << 16)
    u = b[i + 6] | (b[i + 7] << 8) | (b[i + 8] << 16)
    import zlib
    assert len(zlib.decompress(b[i + 9:i + 9 + c])) == u == 800000
    out, _ = rootio.read_tree(path)
    assert out["x"].shape == (100000,) and not out["x"].any()


def test_rejects_what_it_cannot_write(tmp_path):
    with pytest.raises(ValueError):
        rootio.write_tree(tmp_path / "a.root", "T", {"x": np.zeros((3, 2))})
    with pytest.raises(ValueError):
        rootio.write_tree(tmp_path / "a.root", "T", {"x": np.zeros(3), "y": np.zeros(4)})
    with pytest.raises(ValueError):
        rootio.write_tree(tmp_path / "a.root", "T", {"x": np.array([1 + 2j, 3j])})
    (tmp_path / "n.root").write_bytes(b"nope" * 100)
    with pytest.raises(ValueError):
        rootio.read_tree(tmp_path / "n.root")


def test_cross_check_with_uproot_where_installed(tmp_path):
    """The one check this image cannot run (no ROOT, no uproot): files of the own writer opened by uproot and uproot's
    files opened by the own reader.  Runs wherever uproot is installed."""
    uproot = pytest.importorskip("uproot")
    cols = columns(3000)
    cols.pop("flag")                      # uproot writes bool and string branches with its own conventions: numbers only
    names = cols.pop("ParticleName")
    path = tmp_path / "ours.root"
    rootio.write_tree(path, "PhaseSpace", cols, compression=1, basket_entries=1000)
    with uproot.open(path) as f:
        t = f["PhaseSpace"]
        assert t.num_entries == 3000 and set(t.keys()) == set(cols)
        got = t.arrays(library="np")
        for k, v in cols.items():
            assert np.array_equal(got[k], v), k
    with_names = tmp_path / "names.root"
    rootio.write_tree(with_names, "PhaseSpace", {"ParticleName": names, "x": cols["KineticEnergy"]})
    with uproot.open(with_names) as f:
        assert f["PhaseSpace"]["ParticleName"].array(library="np").tolist() == names.tolist()
    theirs = tmp_path / "theirs.root"
    with uproot.recreate(theirs, compression=uproot.ZLIB(1)) as f:
        f["PhaseSpace"] = cols
    out, avail = rootio.read_tree(theirs)
    assert set(avail) == set(cols)
    for k, v in cols.items():
        assert np.array_equal(out[k], v), k


def test_round_trip_property(tmp_path):
    """random schemas: any mix of the supported column types, lengths, basket sizes, compression and layouts"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    dtypes = ["float64", "float32", "int64", "int32", "int16", "int8", "uint8", "uint16", "uint32", "bool", "str"]

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.sampled_from(dtypes), min_size=1, max_size=6), st.integers(0, 400), st.integers(1, 300),
           st.sampled_from([0, 1, 9]), st.booleans(), st.integers(0, 2**31 - 1))
    def check(kinds, n, basket, comp, big, seed):
        r = np.random.default_rng(seed)
        cols = {}
        for i, k in enumerate(kinds):
            if k == "str":
                words = np.array(["", "gamma", "e+", "γ", "y" * 260])
                cols[f"b{i}"] = words[r.integers(0, len(words), n)] if n else np.zeros(0, dtype="U1")
            elif k == "bool":
                cols[f"b{i}"] = r.integers(0, 2, n).astype(bool)
            elif k.startswith("float"):
                cols[f"b{i}"] = r.normal(size=n).astype(k)
            else:
                info = np.iinfo(k)
                cols[f"b{i}"] = r.integers(info.min, info.max, n, dtype=k, endpoint=True)
        path = tmp_path / "p.root"
        rootio.write_tree(path, "t", cols, compression=comp, basket_entries=basket, big=big)
        out, avail = rootio.read_tree(path)
        assert avail == list(cols)
        for name, v in cols.items():
            assert np.array_equal(out[name], v), (name, kinds)
            assert out[name].dtype == v.dtype or v.dtype.kind == "U"

    check()
