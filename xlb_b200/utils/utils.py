"""Dependency-free stand-ins for reference xlb/utils/utils.py:50-208 (matplotlib / pyvista are not required here)."""

import numpy as np


def _to_numpy(a):
    return a.numpy() if hasattr(a, "numpy") else np.asarray(a)


def save_image(fld, timestep=None, prefix=None, **kwargs):
    """Write a 2-D field as a binary PGM image (reference writes a PNG through matplotlib, utils.py:50-90)."""
    fld = _to_numpy(fld)
    if fld.ndim == 3 and fld.shape[0] in (2, 3):
        fld = np.sqrt((fld**2).sum(axis=0))
    if fld.ndim != 2:
        raise ValueError("save_image expects a 2-D field")
    name = (prefix or "field") + ("_" + str(timestep).zfill(7) if timestep is not None else "") + ".pgm"
    lo, hi = float(np.nanmin(fld)), float(np.nanmax(fld))
    img = np.zeros_like(fld, dtype=np.uint8) if hi <= lo else ((fld - lo) / (hi - lo) * 255.0).astype(np.uint8)
    img = np.ascontiguousarray(img.T[::-1])
    with open(name, "wb") as fh:
        fh.write(f"P5 {img.shape[1]} {img.shape[0]} 255\n".encode())
        fh.write(img.tobytes())
    return name


def save_fields_vtk(fields, timestep, output_dir=".", prefix="fields", **kwargs):
    """Write scalar fields as a legacy-VTK structured-points file (reference uses pyvista, utils.py:93-135)."""
    import os

    arrays = {k: _to_numpy(v) for k, v in fields.items()}
    shape = next(iter(arrays.values())).shape
    dims = tuple(shape) + (1,) * (3 - len(shape))
    os.makedirs(output_dir, exist_ok=True)
    name = os.path.join(output_dir, f"{prefix}_{str(timestep).zfill(7)}.vtk")
    with open(name, "w") as fh:
        fh.write("# vtk DataFile Version 3.0\nxlb_b200\nASCII\nDATASET STRUCTURED_POINTS\n")
        fh.write(f"DIMENSIONS {dims[0]} {dims[1]} {dims[2]}\nORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA {int(np.prod(dims))}\n")
        for key, arr in arrays.items():
            fh.write(f"SCALARS {key} float 1\nLOOKUP_TABLE default\n")
            np.savetxt(fh, np.asarray(arr, dtype=np.float32).reshape(dims).transpose(2, 1, 0).reshape(-1), fmt="%.7g")
    return name


def read_stl(path):
    """Triangle soup of an STL file (binary or ASCII) as a float64 array (3 T, 3): three consecutive rows per triangle —
    the layout `mesh_vertices` of a mesh-based boundary condition expects (reference scripts obtain it from
    `trimesh.load_mesh(path, process=False).vertices`, examples/cfd/windtunnel_3d.py:76-78)."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) >= 84:
        n = int(np.frombuffer(data, dtype="<u4", count=1, offset=80)[0])
        if len(data) == 84 + 50 * n:  # binary: 80-byte header, count, 50-byte records (normal, 3 vertices, attribute)
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
            return rec["v"].reshape(-1, 3).astype(np.float64)
    rows = [line.split()[1:4] for line in data.decode("ascii", "replace").splitlines() if line.strip().startswith("vertex")]
    verts = np.array(rows, dtype=np.float64).reshape(-1, 3)
    if verts.shape[0] == 0 or verts.shape[0] % 3:
        raise ValueError(f"{path}: not an STL file (found {verts.shape[0]} vertices)")
    return verts
