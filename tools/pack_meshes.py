#!/usr/bin/env python
"""Pack the reference's triangle-soup DATA files (simulator/*.tris: "ntris" then 9 numbers per triangle,
collisionDetection.cpp:502-518) into the binary mesh file the C-ABI and the oracle load at run time:

    "CKCMESH1" | int32 nmesh | per mesh: int32 id | int32 ntris | ntris*9 float64 (little endian)

The doubles are the strtod values of the text tokens, i.e. exactly what the reference's fscanf("%lf") produces.
Usage: python tools/pack_meshes.py [/root/reference/simulator] [out]
"""
import os
import struct
import sys
import numpy as np

MESHES = ["Base_r", "Link1_r", "Link2_r", "Link3_r", "Link4_r2", "Link5_r", "Link6", "EE_r", "table", "obs", "cone"]


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/simulator"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                               "abb_dualarm_planning_b200", "data", "meshes.ckcmesh")
    blob = b"CKCMESH1" + struct.pack("<i", len(MESHES))
    for mid, name in enumerate(MESHES):
        tok = open(os.path.join(src, name + ".tris")).read().split()
        n = int(tok[0])
        v = np.array([float(t) for t in tok[1:1 + 9 * n]], dtype="<f8")
        assert v.size == 9 * n, name
        blob += struct.pack("<ii", mid, n) + v.tobytes()
        print("%-10s %5d triangles" % (name, n))
    with open(out, "wb") as f:
        f.write(blob)
    print("wrote", out, len(blob), "bytes")


if __name__ == "__main__":
    main()
