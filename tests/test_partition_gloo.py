"""N>1 host logic on CPU: two gloo ranks compute the strip partition the CUDA library uses
(mesh_rectangle in csrc/hdg_mesh.cu, mirrored by hdg.strip_partition) and exchange the 128-byte
communicator id the way `_Context.comm_init` does.  No GPU, no NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import hdg_b200 as hdg
import hdg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import hdg_b200 as hdg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nx, ny = 7, 5
part = hdg.strip_partition(nx, ny, rank, world)
# the unique-id plumbing of _Context.comm_init: rank 0 creates 128 bytes, everybody receives them
buf = np.zeros(128, dtype=np.uint8)
if rank == 0:
    buf[:] = np.arange(128, dtype=np.uint8) ^ 0x5a
t = torch.from_numpy(buf)
dist.broadcast(t, src=0)
parts = [None] * world
dist.all_gather_object(parts, (part, bytes(t.numpy().tobytes())))
if rank == 0:
    print(json.dumps([[p, list(b)] for p, b in parts]))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_partition_and_id_broadcast(tmp_path):
    import json
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(w), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("[[")][-1]
    parts = json.loads(line)
    expect_id = list((np.arange(128, dtype=np.uint8) ^ 0x5a).tolist())
    assert all(b == expect_id for _, b in parts)
    p0, p1 = parts[0][0], parts[1][0]
    nx, ny = 7, 5
    mo = orc.rectangle_mesh(nx, ny)
    # strips tile the global cell and face ranges
    assert p0["cell_begin"] == 0 and p0["cell_end"] == p1["cell_begin"] and p1["cell_end"] == mo.ncells
    assert p0["face_begin"] == 0 and p0["face_end"] == p1["face_begin"] and p1["face_end"] == mo.nfaces
    # every face a rank owns is created (first encountered) by one of its own cells
    for p in (p0, p1):
        cells = np.arange(p["cell_begin"], p["cell_end"])
        created = set((np.flatnonzero(np.isin(mo.faces[:, 2] - 1, cells))).tolist())
        assert created == set(range(p["face_begin"], p["face_end"]))
    # ghost layer sizes: rank 0 sees 2*nx faces + nx cells above, rank 1 sees nx faces below
    assert (p0["ghost_cells"], p0["ghost_faces"]) == (nx, 2 * nx)
    assert (p1["ghost_cells"], p1["ghost_faces"]) == (0, nx)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_tiles_for_any_world(world):
    nx, ny = 5, 11
    parts = [hdg.strip_partition(nx, ny, r, world) for r in range(world)]
    assert parts[0]["cell_begin"] == 0 and parts[-1]["cell_end"] == 2 * nx * ny
    assert parts[0]["face_begin"] == 0 and parts[-1]["face_end"] == 3 * nx * ny + nx + ny
    for a, b in zip(parts[:-1], parts[1:]):
        assert a["cell_end"] == b["cell_begin"] and a["face_end"] == b["face_begin"] and a["j1"] == b["j0"]
    with pytest.raises(ValueError):
        hdg.strip_partition(4, 2, 0, 3)


def test_quad_face_base_matches_oracle():
    nx, ny = 6, 4
    mo = orc.rectangle_mesh(nx, ny)
    for j in range(ny):
        for i in range(nx):
            q = j * nx + i
            assert mo.cell_faces[2 * q, 0] - 1 == hdg.quad_face_base(i, j, nx)   # the diagonal is the first face a quad creates


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
