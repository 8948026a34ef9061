"""Regenerates tests/golden/triangle_meshes.json from the reference's Triangle fixtures
(test/mesh/figure2.1.*, test/mesh/figure.1.* under /root/reference).  Run in the build container only:
/root/reference does not exist on the GPU box, which is why the result is committed.

    python tests/golden/make_fixtures.py
"""
import json
import os
import re

REF = "/root/reference/test/mesh"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "triangle_meshes.json")
NUM = re.compile(r"\b((\d*\.)?\d+)\b")


def rows(path):
    out, first = [], True
    with open(path) as fh:
        for ln in fh:
            if re.match(r"^\s*(?:#|$)", ln):
                continue
            if first:
                first = False
                continue
            out.append([m.group(0) for m in NUM.finditer(ln)])
    return out


def main():
    data = {}
    for name in ("figure2.1", "figure.1"):
        root = os.path.join(REF, name)
        data[name] = {
            "node": [[float(r[1]), float(r[2]), int(r[3])] for r in rows(root + ".node")],      # x y boundary-marker
            "ele": [[int(r[1]), int(r[2]), int(r[3])] for r in rows(root + ".ele")],            # 1-based node ids
            "edge": [[int(r[1]), int(r[2]), int(r[3])] for r in rows(root + ".edge")],          # v1 v2 boundary-marker
        }
    with open(OUT, "w") as fh:
        json.dump(data, fh, separators=(",", ":"))
    print("wrote", OUT, {k: {kk: len(vv) for kk, vv in v.items()} for k, v in data.items()})


if __name__ == "__main__":
    main()
