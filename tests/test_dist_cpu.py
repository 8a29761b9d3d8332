"""CPU tests of the multi-GPU host logic: the block-column-cyclic ownership map (mirrors csrc/gpx_api.cu) and the
torch.distributed plumbing, exercised with the gloo backend at world_size 2 (no GPU needed)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from thesis_code_b200 import distributed as gd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N,W,world", [(1000, 4, 1), (5000, 4, 2), (200000, 4, 8), (9000, 3, 4), (130, 8, 2), (40000, 1, 8)])
def test_ownership_partitions_all_tile_columns(N, W, world):
    nt = gd.num_tiles(N)
    seen = []
    for r in range(world):
        cols = gd.owned_tile_columns(N, W, r, world)
        assert cols == sorted(cols)
        assert all(gd.panel_owner(c // W, world) == r for c in cols)
        seen += cols
    assert sorted(seen) == list(range(nt))                       # every column owned exactly once
    total = sum(gd.local_factor_bytes(N, W, r, world) for r in range(world))
    assert total == nt * (nt + 1) // 2 * 128 * 128 * 8           # shards add up to the packed lower triangle
    if world > 1 and nt >= 8 * W * world:
        sizes = [gd.local_factor_bytes(N, W, r, world) for r in range(world)]
        assert max(sizes) / min(sizes) < 1.25                    # cyclic layout balances storage


def test_c5_shard_fits_a_b200():
    # BASELINE config C5: N=200000 over 2/4/8 GPUs must fit 180 GB per GPU with room for the panel buffers
    for world in (2, 4, 8):
        worst = max(gd.local_factor_bytes(200000, 4, r, world) for r in range(world))
        assert worst < 90e9 if world == 2 else worst < 45e9


def test_test_tile_shards_cover_everything():
    for Ns, world in [(2500, 8), (100000, 4), (1, 2), (128, 3), (129, 2)]:
        T = gd.num_tiles(Ns)
        edges = [gd.test_tile_shard(Ns, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == T
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1


def test_receive_buffers_alternate_over_non_owned_panels():
    for world in (2, 3, 8):
        for rank in range(world):
            last = None
            for p in range(40):
                if gd.panel_owner(p, world) == rank:
                    continue
                idx = gd.recv_buffer_index(p, rank, world)
                assert idx in (0, 1)
                if last is not None:
                    assert idx != last                           # consecutive received panels use different buffers
                last = idx


@pytest.mark.parametrize("nt,world", [(157, 2), (313, 2), (313, 4), (782, 8), (9, 2), (3, 8), (1, 2)])
def test_sharded_gradient_row_blocks_cover_and_balance(nt, world):
    """gpx_lml_grad on a sharded handle: each rank solves one contiguous block of tile rows of L^-T; the blocks cover all
    rows once and, for problems with enough rows, carry the same share of the ~(nt - i)^2 work."""
    b = gd.inverse_row_blocks(nt, world)
    assert b[0] == 0 and b[-1] == nt and len(b) == world + 1 and all(b[i] <= b[i + 1] for i in range(world))
    if nt >= 16 * world:
        cost = [sum((nt - i) ** 2 + 4.0 * nt for i in range(b[r], b[r + 1])) for r in range(world)]
        assert max(cost) / min(cost) < 1.25, (b, cost)
        assert b[1] - b[0] < b[-1] - b[-2]          # the early rows are the expensive ones: the first block is the shortest


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from thesis_code_b200 import distributed as gd
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()

class FakeLib:
    def gpx_comm_unique_id(self, buf):
        buf.raw = bytes(range(128)); return 0
class FakeHandle:
    device = 0; lib = FakeLib(); got = None
    def comm_init(self, r, w, uid): self.got = (r, w, uid)

h = FakeHandle()
gd.init_handle_comm(h, dist)
assert h.got[0] == rank and h.got[1] == world and h.got[2] == bytes(range(128)), h.got
# the column ownership of the two ranks is disjoint and complete (exchanged through gloo)
mine = torch.zeros(gd.num_tiles(5000), dtype=torch.int32)
mine[gd.owned_tile_columns(5000, 4, rank, world)] = 1
dist.all_reduce(mine)
assert bool((mine == 1).all())
# every rank derives the same predict shards
lo, hi = gd.test_tile_shard(2500, rank, world)
t = torch.tensor([hi - lo]); dist.all_reduce(t)
assert int(t) == gd.num_tiles(2500)
# a sharded fit is a collective over identical inputs: the fingerprint check passes on equal data and raises -- on every
# rank, instead of hanging -- when one rank was given something else (data or hyper-parameters)
X = np.linspace(0, 1, 40).reshape(20, 2); Y = np.sin(X[:, :1]); theta = np.array([1.0, 0.5, 0.01])
gd.check_same_problem(h, X, Y, theta, dist)
for bad in ("X", "theta"):
    Xr = X + (1e-9 if (bad == "X" and rank == 1) else 0.0)
    tr = theta * (1.0 + (1e-12 if (bad == "theta" and rank == 1) else 0.0))
    try:
        gd.check_same_problem(h, Xr, Y, tr, dist)
        raise SystemExit("mismatch in %s not detected" % bad)
    except RuntimeError as e:
        assert "different data or hyper-parameters" in str(e), e
# sharded do_optimize / HMC: the ranks re-seed the global numpy RNG from one broadcast seed and then draw the same numbers
np.random.seed(100 + rank)
seed = gd.sync_rng(h, dist)
draw = torch.from_numpy(np.random.normal(size=4)); ref = draw.clone(); dist.broadcast(ref, src=0)
assert seed is not None and bool((draw == ref).all())
# the bench's agreement on the number of timed steps: min over ranks
t = torch.tensor([float(3 + rank)]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
assert int(t) == 3
dist.destroy_process_group()
print("OK", rank)
"""


def test_gloo_world2_comm_bootstrap(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "OK %d" % r in o, o
