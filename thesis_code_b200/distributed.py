"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous, NCCL inside libgpx for the data path.

The factor is 1-D block-column-cyclic: tile columns are grouped into outer panels of ``outer_tiles`` 128-wide columns and
panel ``p`` is owned by rank ``p % world``.  The helpers below restate that ownership map on
the host (they mirror ``panel_owner`` / ``next_owned_panel`` / ``recv_index`` in csrc/gpx_api.cu) so the host logic can be
tested on CPU with gloo, and bootstrap the library's NCCL communicator from a torch.distributed process group.
"""
import numpy as np

TILE = 128


def num_tiles(N):
    return (int(N) + TILE - 1) // TILE


def num_panels(N, outer_tiles):
    return (num_tiles(N) + outer_tiles - 1) // outer_tiles


def panel_owner(p, world):
    return p % world


def owned_panels(N, outer_tiles, rank, world):
    return list(range(rank, num_panels(N, outer_tiles), world))


def owned_tile_columns(N, outer_tiles, rank, world):
    nt = num_tiles(N)
    cols = []
    for p in owned_panels(N, outer_tiles, rank, world):
        cols.extend(range(p * outer_tiles, min((p + 1) * outer_tiles, nt)))
    return cols


def local_factor_bytes(N, outer_tiles, rank, world):
    """Bytes of packed lower-triangular tile storage rank ``rank`` holds (tile column c stores nt - c tiles)."""
    nt = num_tiles(N)
    return sum((nt - c) for c in owned_tile_columns(N, outer_tiles, rank, world)) * TILE * TILE * 8


def test_tile_shard(Ns, rank, world):
    """Contiguous shard [t0, t1) of the ceil(Ns/128) test tiles handled by ``rank`` in gpx_predict."""
    T = num_tiles(Ns)
    return (T * rank) // world, (T * (rank + 1)) // world


def recv_buffer_index(p, rank, world):
    """Which of the two panel receive buffers a non-owner uses for panel p (alternates over non-owned panels)."""
    owned_before = (p - rank + world - 1) // world if p > rank else 0
    return (p - owned_before) & 1


def inverse_row_blocks(nt, world):
    """Boundaries [b_0 = 0, ..., b_world = nt] of the contiguous tile-row blocks of L^-T that the ranks of a sharded
    ``gpx_lml_grad`` solve (mirrors the split in csrc/gpx_api.cu): row i costs ~ (nt - i)^2 + 4 nt tile steps (its far
    updates + in-panel work), the blocks balance that cost."""
    bnd = [0] + [nt] * world
    if world > 1:
        w = [(nt - i) * (nt - i) + 4.0 * nt for i in range(nt)]
        total, cum, r = sum(w), 0.0, 1
        for i in range(nt):
            if r >= world:
                break
            cum += w[i]
            while r < world and cum >= total * r / world:
                bnd[r] = i + 1
                r += 1
    return bnd


def init_handle_comm(handle, dist, rank=None, world=None):
    """Create the NCCL communicator inside ``handle`` from a torch.distributed process group: rank 0 draws the NCCL
    unique id, it is broadcast as bytes through ``dist`` (works with gloo and nccl backends), every rank joins."""
    import ctypes

    import torch
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1:
        handle.comm_init(0, 1, None)
        return
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = handle.lib.gpx_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("gpx_comm_unique_id failed with code {} (is libnccl.so.2 loadable?)".format(rc))
    device = torch.device("cuda", handle.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    uid = bytes(t.cpu().numpy().astype(np.uint8).tolist())
    handle.comm_init(rank, world, uid)


def problem_fingerprint(X, Y, theta):
    """Six numbers that must agree on every rank of a sharded fit: shapes and CRC32s of the data and the parameters."""
    import zlib

    def crc(a):
        return float(zlib.crc32(np.ascontiguousarray(a, dtype=np.float64).view(np.uint8).ravel()))
    return np.array([X.shape[0], X.shape[1], Y.shape[1], crc(X), crc(Y), crc(theta)], dtype=np.float64)


def check_same_problem(handle, X, Y, theta, dist=None):
    """A sharded fit is a collective: every rank must pass the same X, Y and hyper-parameters, otherwise panels of
    different problems get mixed (or the ranks hang).  One small all-reduce (max of [f, -f]) compares the fingerprints."""
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    fp = problem_fingerprint(X, Y, theta)
    device = torch.device("cuda", handle.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.concatenate([fp, -fp])).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    if not np.array_equal(t[:fp.size], -t[fp.size:]):
        from ._gpx import GpxError
        raise GpxError("sharded GPModel: the ranks were given different data or hyper-parameters "
                       "(N, d, O, crc(X), crc(Y), crc(theta) max {} vs min {}); a sharded fit is a collective over "
                       "identical inputs -- use model.distributed = False for per-rank models".format(t[:fp.size], -t[fp.size:]))


def sync_rng(handle, dist=None):
    """Sharded ``do_optimize`` / HMC: every rank has to walk the same path (each objective evaluation is a collective fit
    over identical hyper-parameters), but GPy's recipe draws its start point and its HMC momenta from the process-global
    numpy RNG.  Rank 0 draws a seed from that RNG, it is broadcast, and every rank re-seeds ``np.random`` with it.
    Returns the seed (None when not distributed)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    device = torch.device("cuda", handle.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(np.random.randint(0, 2 ** 31 - 1))], dtype=torch.int64, device=device)
    dist.broadcast(t, src=0)
    seed = int(t.item())
    np.random.seed(seed)
    return seed
