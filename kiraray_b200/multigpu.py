"""Work partition of the WavefrontPathTracer pass across the GPUs of one box (one process per GPU).

The reference is single-device (src/core/device/context.cpp:37-40); SURVEY.md section 8(e) derives
the split from the path's own structure: pixels are independent (per-pixel RNG stream and
accumulator, integrator.cpp:213-220) and frames are independent samples (the reference averages
frames in AccumulatePass, accumulate.cu:30-52).  So the work is a T x S grid of
  * T image tiles  -- contiguous row bands; a rank renders only its rows and writes zeros elsewhere
                      (krr_wfpt_set_partition), so tile films ADD to the full film, and
  * S spp slices   -- rank s renders frame indices  first + s, first + s + S, ...  (every frame index
                      seeds a different PCG sequence: sampleIndex = frameIndex * spp),
with the scene replicated.  The ONE exchange step is the film accumulation: a sum-reduce of the
RGBA32F film to rank 0, then a division by S.  On the GPUs that is the PRODUCT's own entry point
(krr_wfpt_reduce_film: ncclReduce over NVLink inside libkrr_wfpt.so, include/krr_wfpt.h; the C++ host
layer drives the same entry from one process with one thread per device, host/multi_device.cpp);
this module only carries the partition arithmetic and hands the NCCL unique id from rank 0 to the
other processes.  `reduce_film` below (torch.distributed) is the CPU stand-in of the gloo tests.
No other collective is on the path.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Partition:
    rank: int
    world: int
    tiles: int        # T
    spp_slices: int   # S;  T * S == world
    height: int

    @property
    def tile(self):
        return self.rank % self.tiles

    @property
    def spp_slice(self):
        return self.rank // self.tiles

    @property
    def rows(self):
        """[begin, end) rows of this rank's tile; bands differ by at most one row."""
        t, T, H = self.tile, self.tiles, self.height
        return (t * H) // T, ((t + 1) * H) // T

    def frame_index(self, step, first_frame=1, batch=1):
        """First frame index this rank renders at `step` (first frame is 1, src/core/window.cpp:457); with a frame
        batch every step of a rank covers `batch` consecutive frame indices."""
        return first_frame + (self.spp_slice + step * self.spp_slices) * batch

    def describe(self):
        return f"tile {self.tiles} x spp-by-frame {self.spp_slices}, scene replicated, film sum-reduce to rank 0"


def make_partition(rank, world, height, mode="spp", tiles=None):
    """mode 'spp': S = world; 'tile': T = world; 'hybrid': T = tiles (must divide world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if mode == "spp":
        T = 1
    elif mode == "tile":
        T = world
    elif mode == "hybrid":
        T = tiles or (2 if world % 2 == 0 and world > 1 else 1)
    else:
        raise ValueError(f"unknown partition mode {mode!r}")
    if world % T != 0:
        raise ValueError(f"tiles={T} does not divide world={world}")
    if T > height:
        raise ValueError("more tiles than rows")
    return Partition(rank, world, T, world // T, height)


def reduce_film(film, part, dist=None, dst=0):
    """Film accumulation: sum over ranks to `dst`, then the average over the spp slices.  `film` is a
    torch tensor (H, W, 4) -- CUDA for NCCL, CPU for gloo; modified in place on dst."""
    if dist is not None and part.world > 1:
        dist.reduce(film, dst=dst)
    if part.rank == dst and part.spp_slices > 1:
        film /= part.spp_slices
    return film


class FilmReducer:
    """One process per GPU (torchrun): binds the pass handle of this rank to the film-reduction communicator of
    the product library.  The NCCL unique id is created by rank 0 (krr_wfpt_comm_unique_id) and travels to the
    other processes through the torch.distributed process group -- the only thing torch is used for here."""

    def __init__(self, gpu, part, dist=None, use_torch=False):
        """use_torch: sum the films with torch.distributed.reduce (reduce_film above) instead of the library's own NCCL
        entry points -- a diagnostic switch (bench.py --reduce torch), not the product path."""
        self.gpu, self.part, self.dist = gpu, part, dist
        self.scale = 1.0 / part.spp_slices
        self.native = False
        self.use_torch = bool(use_torch) and dist is not None and part.world > 1
        if dist is not None and part.world > 1 and not self.use_torch:
            import torch
            # every rank checks that the library can reach NCCL at all (krr_wfpt_comm_unique_id dlopens libnccl.so.2); the
            # ranks must take the same path, so the answers are combined first.  Without NCCL in the library the films
            # are still summed -- through torch.distributed -- and the fact is reported (self.native stays False).
            my_uid, ok = None, 1
            try:
                my_uid = gpu.comm_unique_id()
            except RuntimeError as e:
                ok, self.fallback_reason = 0, str(e)
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"  # (gloo: the CPU tests of this negotiation)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 0:
                self.use_torch = True
                return
            uid = torch.zeros(128, dtype=torch.uint8)
            if part.rank == 0:
                uid = torch.frombuffer(bytearray(my_uid), dtype=torch.uint8).clone()
            uid = uid.to(dev)
            dist.broadcast(uid, 0)
            gpu.comm_init_rank(bytes(uid.cpu().numpy().tobytes()), part.world, part.rank)
            self.native = True

    def reduce(self, film, stream=None):
        """film: CUDA tensor (H, W, 4); summed onto rank 0 in place and divided by the spp slices there."""
        if self.use_torch:
            reduce_film(film, self.part, self.dist)
        elif self.native or self.scale != 1.0:
            self.gpu.reduce_film(film.data_ptr(), 0, self.scale, stream)

    def render_reduce_to_host_async(self, film_host, stream=None):
        if self.use_torch:  # no pipelining on this path: render, reduce, synchronous copy on rank 0
            import torch
            if not hasattr(self, "_film"):
                self._film = torch.empty((self.part.height, self.gpu.size[0], 4), dtype=torch.float32, device="cuda")
            self.gpu.render(self._film.data_ptr(), stream)
            reduce_film(self._film, self.part, self.dist)
            if self.part.rank == 0:
                torch.from_numpy(film_host).copy_(self._film)
            return
        self.gpu.render_reduce_to_host_async(film_host if self.part.rank == 0 else None, 0, self.scale, stream)

    def wait_host(self):
        self.gpu.wait_host()

    def close(self):
        if self.native:
            self.gpu.comm_destroy()
            self.native = False
