"""Sample clouds and a minimal PLY vertex reader
(reference: or_pcd/data/__init__.py:5-19 uses pkg_resources + o3d.io.read_point_cloud).

The Stanford Armadillo range scans the reference ships as big-endian PLY files are
kept here, inside the package, as one compressed ``.npz`` of their float32 vertex arrays
(``armadillo_scans.npz`` next to this file); Open3D's reader widens to float64, so do we."""
import os
import struct

import numpy as np

_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "armadillo_scans.npz")

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2",
              "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def sample_cloud_names():
    with np.load(_NPZ) as z:
        return sorted(z.files)


def load_sample_cloud(name: str = "ArmadilloBack_180") -> np.ndarray:
    """(N,3) float64 vertices of one of the bundled scans (reference signature: data/__init__.py:5,
    ``load_sample_cloud(name="ArmadilloBack_180")``; it returns an Open3D cloud there, the vertex array here)."""
    with np.load(_NPZ) as z:
        if name not in z.files:
            raise FileNotFoundError(f"unknown sample cloud {name!r}; have {sorted(z.files)}")
        return z[name].astype(np.float64)


def load_cloud(path: str) -> np.ndarray:
    """(N,3) float64 cloud from a file, the two on-disk formats of the reference's driver
    (main.py:55-71): ``.npz`` archives with a ``points`` array (first three columns), anything else
    is read as PLY (``o3d.io.read_point_cloud`` in the reference)."""
    if path.lower().endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:   # a numeric array needs no pickle; object arrays are refused
            if "points" not in z.files:
                raise ValueError(f"{path}: no 'points' array in the archive")
            try:
                pts = z["points"]
            except ValueError as e:
                raise ValueError(f"{path}: 'points' must be a plain numeric array (object arrays are not loaded)") from e
            if pts.ndim != 2 or pts.shape[1] < 3 or not np.issubdtype(pts.dtype, np.number):
                raise ValueError(f"{path}: 'points' must be a numeric (N, >=3) array, got {pts.dtype} {pts.shape}")
            return np.ascontiguousarray(pts[:, :3], dtype=np.float64)
    return read_ply_vertices(path)


def read_ply_vertices(path: str) -> np.ndarray:
    """x,y,z of the ``vertex`` element of an ascii / binary_little_endian / binary_big_endian
    PLY file as (N,3) float64.  Elements before ``vertex`` must be fixed-size."""
    with open(path, "rb") as f:
        buf = f.read()
    end = buf.index(b"end_header\n") + len(b"end_header\n")
    header = buf[:end].decode("ascii", errors="replace").splitlines()
    if header[0].strip() != "ply":
        raise ValueError("not a PLY file")
    fmt = None
    elements = []   # (name, count, [(prop, type)])
    for line in header[1:]:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], "list"))
            else:
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
    offset = end
    for name, count, props in elements:
        if name == "vertex":
            names = [p[0] for p in props]
            if fmt == "ascii":
                rows = buf[offset:].decode("ascii").split("\n")[:count]
                arr = np.array([[float(v) for v in r.split()] for r in rows], dtype=np.float64)
                return np.ascontiguousarray(arr[:, [names.index("x"), names.index("y"), names.index("z")]])
            order = ">" if fmt == "binary_big_endian" else "<"
            dt = np.dtype([(n, order + t) for n, t in props])
            v = np.frombuffer(buf, dtype=dt, count=count, offset=offset)
            return np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float64)
        if any(t == "list" for _, t in props) or fmt == "ascii":
            raise ValueError("variable-size element before 'vertex' is not supported")
        offset += count * sum(np.dtype(t).itemsize for _, t in props)
    raise ValueError("no vertex element")
