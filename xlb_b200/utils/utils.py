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
