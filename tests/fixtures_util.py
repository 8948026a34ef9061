"""Writes the Triangle (.node/.ele/.edge) fixtures from tests/golden/triangle_meshes.json into a temporary
directory (the JSON is generated from the reference's test meshes by tests/golden/make_fixtures.py)."""
import json
import os
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = None


def triangle_root(name):
    """Path prefix `<tmp>/<name>` such that `<prefix>.node`, `.ele`, `.edge` exist (Shewchuk Triangle format)."""
    global _DIR
    if _DIR is None:
        _DIR = tempfile.mkdtemp(prefix="hdg_triangle_")
        with open(os.path.join(_HERE, "golden", "triangle_meshes.json")) as fh:
            data = json.load(fh)
        for nm, m in data.items():
            root = os.path.join(_DIR, nm)
            with open(root + ".node", "w") as f:
                f.write(f"{len(m['node'])}  2  0  1\n")
                for i, (x, y, mk) in enumerate(m["node"], 1):
                    f.write(f"   {i}    {x!r}  {y!r}    {mk}\n")
                f.write("# written by tests/fixtures_util.py\n")
            with open(root + ".ele", "w") as f:
                f.write(f"{len(m['ele'])}  3  0\n")
                for i, (a, b, c) in enumerate(m["ele"], 1):
                    f.write(f"   {i}    {a}  {b}  {c}\n")
            with open(root + ".edge", "w") as f:
                f.write(f"{len(m['edge'])}  1\n")
                for i, (a, b, mk) in enumerate(m["edge"], 1):
                    f.write(f"   {i}   {a}  {b}  {mk}\n")
    return os.path.join(_DIR, name)
