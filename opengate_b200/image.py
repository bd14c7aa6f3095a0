"""Minimal image container + MetaImage I/O + the image helpers on the hot path.

The reference passes `itk.Image` objects around (opengate/image.py); ITK is a binary
dependency that the GPU engine does not need.  `Image` keeps what the path uses: a numpy array
in (z,y,x) order (the reference's numpy view of an itk image, helpers_itk_image_py.h:171-257),
spacing, origin and direction in (x,y,z) order, and reads/writes MetaImage (.mhd + .raw),
the format the reference's tests and writers use (opengate/actors/dataitems.py:1336-1340).
If `itk` is importable, `to_itk()/from_itk()` convert.
"""

from __future__ import annotations

import os

import numpy as np

from .exception import fatal

_MET_TYPES = {
    "MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16,
    "MET_UINT": np.uint32, "MET_INT": np.int32, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64,
    "MET_ULONG": np.uint64, "MET_LONG": np.int64,
}
_NP_TO_MET = {np.dtype(v): k for k, v in _MET_TYPES.items()}


class Image:
    """3-D image: array (z,y,x), spacing/origin (x,y,z), direction 3x3."""

    def __init__(self, array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=None):
        self.array = np.ascontiguousarray(array)
        if self.array.ndim != 3:
            fatal(f"Image needs a 3-D array, got shape {self.array.shape}")
        self.spacing = np.asarray(spacing, dtype=np.float64).copy()
        self.origin = np.asarray(origin, dtype=np.float64).copy()
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3).copy()

    @property
    def size(self):
        """(nx, ny, nz) like itk's GetLargestPossibleRegion().GetSize()."""
        return np.array(self.array.shape[::-1], dtype=np.int64)

    def copy_information(self, other):
        self.spacing = other.spacing.copy()
        self.origin = other.origin.copy()
        self.direction = other.direction.copy()

    def index_to_physical_point(self, index_xyz):
        return self.origin + self.direction @ (np.asarray(index_xyz, float) * self.spacing)

    def physical_point_to_index(self, p):
        c = (self.direction.T @ (np.asarray(p, float) - self.origin)) / self.spacing
        idx = np.floor(c + 0.5).astype(np.int64)
        inside = bool(np.all(idx >= 0) and np.all(idx < self.size))
        return idx, inside

    def __repr__(self):
        return f"Image(size={self.size.tolist()}, spacing={self.spacing.tolist()}, origin={self.origin.tolist()}, dtype={self.array.dtype})"


def create_3d_image(size, spacing, origin=None, pixel_type="float", allocate=True, fill_value=0):
    """As opengate/image.py:40-59 (size, spacing in x,y,z order)."""
    dt = {"float": np.float32, "double": np.float64, "unsigned short": np.uint16, "short": np.int16,
          "unsigned char": np.uint8, "int": np.int32}.get(pixel_type, pixel_type)
    nx, ny, nz = [int(s) for s in size]
    arr = np.full((nz, ny, nx), fill_value, dtype=dt) if allocate else np.empty((nz, ny, nx), dtype=dt)
    return Image(arr, spacing, (0, 0, 0) if origin is None else origin)


def read_image(path):
    """Read .mhd/.mha (raw or zlib-compressed element data, local or in a separate .raw / .zraw file) or .npy."""
    path = str(path)
    if path.endswith(".npy"):
        return Image(np.load(path))
    if not (path.endswith(".mhd") or path.endswith(".mha")):
        fatal(f"Unsupported image format (need .mhd/.mha): {path}")
    hdr = {}
    offset = 0
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            txt = line.decode("latin-1").strip()
            if "=" not in txt:
                continue
            k, v = [t.strip() for t in txt.split("=", 1)]
            hdr[k] = v
            if k == "ElementDataFile":
                offset = f.tell()
                break
    compressed = hdr.get("CompressedData", "False") == "True"   # ITK: the element data as ONE zlib stream (.zraw)
    ndim = int(hdr.get("NDims", 3))
    dims = [int(x) for x in hdr["DimSize"].split()]
    if ndim != 3:
        fatal(f"Only 3-D MetaImage supported, NDims={ndim}")
    dt = np.dtype(_MET_TYPES[hdr["ElementType"]])
    if hdr.get("BinaryDataByteOrderMSB", hdr.get("ElementByteOrderMSB", "False")) == "True":
        dt = dt.newbyteorder(">")
    n = dims[0] * dims[1] * dims[2]
    if hdr["ElementDataFile"] == "LOCAL":
        with open(path, "rb") as f:
            f.seek(offset)
            blob = f.read() if compressed else f.read(n * dt.itemsize)
    else:
        raw = os.path.join(os.path.dirname(path), hdr["ElementDataFile"])
        with open(raw, "rb") as f:
            blob = f.read() if compressed else f.read(n * dt.itemsize)
    if compressed:
        import zlib
        csize = int(hdr.get("CompressedDataSize", len(blob)))
        try:
            blob = zlib.decompress(blob[:csize])
        except zlib.error as why:
            fatal(f"Cannot decompress the element data of {path}: {why}")
    if len(blob) < n * dt.itemsize:
        fatal(f"{path}: element data hold {len(blob)} bytes, the header asks for {n * dt.itemsize}")
    data = np.frombuffer(blob, dtype=dt, count=n)
    arr = data.reshape(dims[2], dims[1], dims[0]).astype(dt.newbyteorder("="), copy=False)
    spacing = [float(x) for x in hdr.get("ElementSpacing", "1 1 1").split()]
    origin = [float(x) for x in hdr.get("Offset", hdr.get("Position", "0 0 0")).split()]
    tm = [float(x) for x in hdr.get("TransformMatrix", "1 0 0 0 1 0 0 0 1").split()]
    # MetaImage stores the direction cosines column by column
    direction = np.array(tm).reshape(3, 3).T
    return Image(arr, spacing, origin, direction)


def write_image(img, path, compressed=False):
    """Write .mhd + .raw (or .mha with local data), or .npy; compressed: element data as one zlib stream
    (.mhd + .zraw, CompressedData = True, as ITK's MetaImageIO writes it)."""
    path = str(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if path.endswith(".npy"):
        np.save(path, img.array)
        return
    arr = np.ascontiguousarray(img.array)
    if arr.dtype not in _NP_TO_MET:
        fatal(f"Cannot write dtype {arr.dtype} as MetaImage")
    local = path.endswith(".mha")
    raw_name = "LOCAL" if local else os.path.basename(path)[:-4] + (".zraw" if compressed else ".raw")
    payload = arr.tobytes()
    if compressed:
        import zlib
        payload = zlib.compress(payload, 2)
    nx, ny, nz = img.size
    tm = img.direction.T.reshape(-1)
    hdr = (
        "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\n"
        f"CompressedData = {'True' if compressed else 'False'}\n"
        + (f"CompressedDataSize = {len(payload)}\n" if compressed else "") +
        f"TransformMatrix = {' '.join(repr(float(x)) for x in tm)}\n"
        f"Offset = {' '.join(repr(float(x)) for x in img.origin)}\n"
        "CenterOfRotation = 0 0 0\nAnatomicalOrientation = RAI\n"
        f"ElementSpacing = {' '.join(repr(float(x)) for x in img.spacing)}\n"
        f"DimSize = {nx} {ny} {nz}\n"
        f"ElementType = {_NP_TO_MET[arr.dtype]}\n"
        f"ElementDataFile = {raw_name}\n"
    )
    with open(path, "wb") as f:
        f.write(hdr.encode("latin-1"))
        if local:
            f.write(payload)
    if not local:
        with open(os.path.join(os.path.dirname(os.path.abspath(path)), raw_name), "wb") as f:
            f.write(payload)


def compute_image_3D_CDF(image_or_array):
    """Cumulative distribution functions of a 3-D activity image, numpy order (z,y,x):
    returns (cdf_x [nz][ny][nx], cdf_y [nz][ny], cdf_z [nz]) exactly as the reference's
    opengate/image.py:313-358 (same normalisations; rows that sum to zero stay zero)."""
    array = image_or_array.array if isinstance(image_or_array, Image) else np.asarray(image_or_array)
    array = array / np.sum(array)
    sumx = np.sum(array, axis=2)
    sumxy = np.sum(sumx, axis=1)
    cdf_x = np.cumsum(array, axis=2)
    last = cdf_x[:, :, -1:]
    cdf_x = np.divide(cdf_x, last, out=cdf_x.copy(), where=last != 0)
    cdf_y = np.cumsum(sumx, axis=1)
    last = cdf_y[:, -1:]
    cdf_y = np.divide(cdf_y, last, out=cdf_y.copy(), where=last != 0)
    cdf_z = np.cumsum(sumxy) / np.sum(sumxy)
    return cdf_x, cdf_y, cdf_z


def images_have_same_domain(a, b, tol=1e-5):
    return (np.all(a.size == b.size) and np.allclose(a.spacing, b.spacing, atol=tol)
            and np.allclose(a.origin, b.origin, atol=tol) and np.allclose(a.direction, b.direction, atol=tol))


def to_itk(img):  # pragma: no cover - only where itk exists
    import itk
    out = itk.image_from_array(img.array)
    out.SetSpacing([float(x) for x in img.spacing])
    out.SetOrigin([float(x) for x in img.origin])
    out.SetDirection(itk.matrix_from_array(img.direction))
    return out


def from_itk(itk_image):  # pragma: no cover - only where itk exists
    import itk
    return Image(itk.array_from_image(itk_image), list(itk_image.GetSpacing()), list(itk_image.GetOrigin()),
                 itk.array_from_matrix(itk_image.GetDirection()))
