"""N > 1 host logic on CPU: world_size-2 (and 4) gloo runs of the tile / spp partition and the film
reduce, with the CPU oracle standing in for the renderer (it honours row bands and frame indices the
same way the C ABI does).  The tile split must reproduce the single-process film bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kiraray_b200.multigpu import make_partition, reduce_film

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = H = 24


def test_partition_covers_every_row_once_and_frames_are_disjoint():
    for world in (1, 2, 3, 4, 8):
        for mode in ("spp", "tile", "hybrid"):
            if mode == "hybrid" and world % 2:
                continue
            parts = [make_partition(r, world, 1080, mode) for r in range(world)]
            for s in range(parts[0].spp_slices):
                rows = sorted(p.rows for p in parts if p.spp_slice == s)
                assert rows[0][0] == 0 and rows[-1][1] == 1080
                assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            frames = {(p.tile, p.frame_index(k)) for p in parts for k in range(5)}
            assert len(frames) == world * 5
            assert sorted({p.frame_index(k) for p in parts for k in range(5)}) == list(range(1, 1 + 5 * parts[0].spp_slices))
    with pytest.raises(ValueError):
        make_partition(0, 4, 1080, "hybrid", tiles=3)
    with pytest.raises(ValueError):
        make_partition(4, 4, 1080)


def _render(part, step=0):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import kiraray_b200 as krr
    import oracle_binding as ob
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    app.set_resolution(W, H)
    kind = "reference"
    orc = ob.Oracle(app.scene_desc(), kind)
    out = orc.render(app.camera(), W, H, frame_index=part.frame_index(step), spp=1, max_depth=3, threads=1, rows=part.rows)
    orc.close()
    return out["film"]


def _worker(rank, world, mode, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = make_partition(rank, world, H, mode)
    film = torch.from_numpy(_render(part).copy())
    reduce_film(film, part, dist)
    if rank == 0:
        ret.put(film.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def _run(world, mode, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, mode, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    film = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return film


def test_tile_split_world2_reproduces_the_single_process_film_bit_exactly():
    full = _render(make_partition(0, 1, H))
    got = _run(2, "tile", 29511)
    assert np.array_equal(got.view(np.uint32), full.view(np.uint32))


def test_spp_split_world2_averages_two_frames():
    f1 = _render(make_partition(0, 2, H, "spp"))
    f2 = _render(make_partition(1, 2, H, "spp"))
    got = _run(2, "spp", 29512)
    assert np.allclose(got, (f1 + f2) / 2, rtol=1e-6, atol=0)
    assert not np.array_equal(f1, f2)  # different frame index -> different RNG streams


def test_hybrid_world4_two_tiles_by_two_frames():
    one = make_partition(0, 1, H)
    f1 = _render(one, 0)
    f2 = _render(one, 1)  # frame index 2
    got = _run(4, "hybrid", 29513)
    assert np.allclose(got, (f1 + f2) / 2, rtol=1e-6, atol=0)


class _StubPass:
    """stands in for the pass handle: records what FilmReducer asks of it; `nccl` = whether its library reaches NCCL"""

    def __init__(self, nccl):
        self.nccl, self.calls, self.size = nccl, [], (W, H)

    def comm_unique_id(self):
        self.calls.append("unique_id")
        if not self.nccl:
            raise RuntimeError("krr_wfpt_comm_unique_id: libnccl.so.2 not found")
        return bytes(range(128))

    def comm_init_rank(self, uid, world, rank):
        self.calls.append(("init_rank", uid, world, rank))

    def reduce_film(self, ptr, root, scale, stream):
        self.calls.append(("reduce_film", root, scale))


def _negotiate(rank, world, missing_on, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kiraray_b200.multigpu import FilmReducer
    part = make_partition(rank, world, H, "spp")
    gpu = _StubPass(nccl=rank != missing_on)
    red = FilmReducer(gpu, part, dist)
    film = torch.full((H, W, 4), float(rank + 1))
    red.reduce(film, None)
    ret.put((rank, red.native, red.use_torch, [c if isinstance(c, str) else c[0] for c in gpu.calls],
             [c for c in gpu.calls if not isinstance(c, str) and c[0] == "init_rank"], float(film[0, 0, 0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("missing_on", [-1, 1])
def test_ranks_agree_on_the_film_reduction_path(missing_on):
    """FilmReducer: every rank probes its library for NCCL, the answers are combined, and ALL ranks take the same path --
    the library's communicator (rank 0's id reaches every rank) or, if any rank lacks NCCL, torch.distributed on all of
    them (a rank left alone in a collective would hang the job)."""
    world, port = 2, 29650 + (os.getpid() + missing_on) % 200
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_negotiate, args=(r, world, missing_on, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    out = sorted(ret.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for rank, native, use_torch, calls, inits, v in out:
        if missing_on < 0:
            assert native and not use_torch and calls == ["unique_id", "init_rank", "reduce_film"]
            assert inits[0][1] == bytes(range(128)) and inits[0][2:] == (world, rank)   # rank 0's id, on every rank
        else:
            assert not native and use_torch and calls == ["unique_id"], (rank, calls)
    if missing_on >= 0:   # the torch path summed the films onto rank 0 and averaged over the spp slices
        assert out[0][5] == (1.0 + 2.0) / world
