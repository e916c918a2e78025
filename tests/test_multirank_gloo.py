"""N>1 host path on CPU (gloo, world_size 2): the static interleaved tile split of the frame and
the host gather bench.py performs (reduce-SUM of zero-padded, disjoint per-rank buffers).
The renderer itself needs a GPU; here each rank fills the tiles it owns with a deterministic
function of the pixel coordinates, exactly where libtpt.so would write (tile t belongs to rank
t % world; tiles are TPT_TILE x TPT_TILE, row-major), and rank 0 must end up with the full frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TILE = 16


def owner_map(nx, ny, world):
    tx = (nx + TILE - 1) // TILE
    jj, ii = np.mgrid[0:ny, 0:nx]
    tile = (jj // TILE) * tx + (ii // TILE)
    return tile % world


def fake_render(nx, ny, rank, world):
    """what a rank's tpt_render leaves in its host buffers: its own tiles, zeros elsewhere"""
    jj, ii = np.mgrid[0:ny, 0:nx]
    full = np.stack([ii * 0.5 + jj, ii - 0.25 * jj, (ii * jj) % 7], axis=-1).astype(np.float32)
    mine = owner_map(nx, ny, world) == rank
    return np.where(mine[..., None], full, 0).astype(np.float32), full


def worker(rank, world, port, nx, ny, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part, full = fake_render(nx, ny, rank, world)
    t = torch.from_numpy(part.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    paths = torch.tensor([float((owner_map(nx, ny, world) == rank).sum())], dtype=torch.float64)
    dist.all_reduce(paths, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), t.numpy())
        np.save(os.path.join(out_dir, "meta.npy"), np.array([paths.item(), tmax.item()]))
    dist.barrier()
    dist.destroy_process_group()


def worker_owned(rank, world, port, nx, ny, out_dir):
    """bench.py's multi-rank e2e gather: each rank sends only the pixels of the tiles it owns"""
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part, full = fake_render(nx, ny, rank, world)
    for dtype in (np.float32, np.uint8):
        frame = torch.from_numpy(part.astype(dtype).reshape(nx * ny, 3).copy())
        bench.gather_owned_tiles(torch, dist, frame, nx, ny, rank, world)
        if rank == 0:
            np.save(os.path.join(out_dir, f"owned_{np.dtype(dtype).name}.npy"), frame.numpy().reshape(ny, nx, 3))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nx,ny", [(1200, 1200), (150, 90)])
def test_two_rank_gather_reproduces_the_frame(tmp_path, nx, ny):
    world = 2
    mp.spawn(worker, args=(world, free_port(), nx, ny, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    _, full = fake_render(nx, ny, 0, world)
    assert np.array_equal(got, full)  # disjoint tiles + zeros: the sum is exact
    paths, tmax = np.load(tmp_path / "meta.npy")
    assert paths == nx * ny and tmax == 2.0  # whole-job unit count, max-over-ranks time


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_static_split_is_balanced_and_complete(world):
    """every pixel has exactly one owner; on the headline frame the interleaved split gives every
    rank the same number of tiles to within one, and the same share of the box interior (the
    expensive pixels) to within 3 %."""
    nx = ny = 1200
    own = owner_map(nx, ny, world)
    counts = np.bincount(own.ravel(), minlength=world)
    assert counts.sum() == nx * ny
    assert counts.max() - counts.min() <= TILE * TILE
    jj, ii = np.mgrid[0:ny, 0:nx]
    inside = (np.abs(ii - 600) < 225) & (np.abs(jj - 600) < 225)  # fov 90: the box opening covers ~37.5 % of each axis
    share = np.array([(inside & (own == r)).sum() for r in range(world)], float)
    assert share.max() <= 1.03 * share.mean()


@pytest.mark.parametrize("nx,ny,world", [(1200, 1200, 2), (150, 90, 2), (100, 70, 3)])
def test_tile_owned_gather_reproduces_the_frame(tmp_path, nx, ny, world):
    """the gather bench.py uses under torchrun: 1/world of the frame per rank instead of a full-frame SUM;
    ragged frames (partial tiles, unequal tile counts per rank) included"""
    mp.spawn(worker_owned, args=(world, free_port(), nx, ny, str(tmp_path)), nprocs=world, join=True)
    _, full = fake_render(nx, ny, 0, world)
    assert np.array_equal(np.load(tmp_path / "owned_float32.npy"), full)
    assert np.array_equal(np.load(tmp_path / "owned_uint8.npy"), full.astype(np.uint8))


def test_owned_pixel_index_matches_the_kernel_split():
    """bench.owned_pixel_index == the tile ownership rule of the kernels (tile t belongs to part t % parts)"""
    import bench
    for nx, ny, world in [(1200, 1200, 8), (150, 90, 3)]:
        own = owner_map(nx, ny, world).ravel()
        seen = np.zeros(nx * ny, np.int32)
        for r in range(world):
            idx = bench.owned_pixel_index(torch, nx, ny, r, world, "cpu").numpy()
            assert (own[idx] == r).all() and len(idx) == (own == r).sum()
            seen[idx] += 1
        assert (seen == 1).all()
