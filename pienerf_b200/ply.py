"""Minimal PLY vertex reader/writer for the reference's point-cloud schema (solver.py:109-135,
main_sample.py:14-23: x,y,z (f8), vp, pin, lam, mu, mass).  plyfile is not in this image."""
import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def read_ply_vertices(path):
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError("not a ply file")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = fh.readline()
            if not line:
                raise ValueError("unterminated ply header")
            tok = line.decode("ascii").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(fh, max_rows=n, ndmin=2)
            return {name: data[:, i] for i, (name, _) in enumerate(props)}
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t) for name, t in props])
        arr = np.frombuffer(fh.read(n * dt.itemsize), dtype=dt, count=n)
        return {name: np.asarray(arr[name]) for name, _ in props}


def write_ply_xyz(path, xyz, extra=None):
    """binary little-endian; xyz float64 like Simulator.OutputToPly; extra: dict name -> array (f8)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    cols = [("x", xyz[:, 0]), ("y", xyz[:, 1]), ("z", xyz[:, 2])] + [(k, np.asarray(v, dtype=np.float64)) for k, v in (extra or {}).items()]
    dt = np.dtype([(k, "<f8") for k, _ in cols])
    arr = np.empty(xyz.shape[0], dtype=dt)
    for k, v in cols:
        arr[k] = v
    with open(path, "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\n")
        fh.write(f"element vertex {xyz.shape[0]}\n".encode())
        for k, _ in cols:
            fh.write(f"property double {k}\n".encode())
        fh.write(b"end_header\n")
        fh.write(arr.tobytes())
