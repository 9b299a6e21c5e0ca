"""Multi-GPU host logic: one process per GPU, the frame cut into horizontal strips of whole 16-row tiles.

Every rank runs the SAME pass list on its own strip (`nrd.CudaContext(strip=..., strip_height=...)`); rows of other strips
are read by the kernels straight from the owner's HBM over NVLink (CUDA IPC mapping of the owner's arena), and a device-side
flag barrier separates the passes (csrc/executor.cu).  There is no data-path collective: torch.distributed only carries the
64-byte IPC handles at start-up and the timing reductions of bench.py.
"""
import os

import torch

from . import harness, nrd

TILE = 16
# ghost rows each strip keeps of its neighbours (refreshed by bulk NVLink stores after every pass); taps beyond them are
# direct peer loads, so the value only trades halo traffic against the (slow) direct loads
DEFAULT_HALO_ROWS = int(os.environ.get("NRD_B200_HALO_ROWS", "96"))


def partition_rows(height, world_size):
    """Uniform strips of whole 16-row tiles: returns (strip_height, [(y0, y1)] per rank).  Every rank must own rows."""
    tiles = (height + TILE - 1) // TILE
    per_rank = (tiles + world_size - 1) // world_size
    strip_height = per_rank * TILE
    strips = [(r * strip_height, min((r + 1) * strip_height, height)) for r in range(world_size)]
    if any(y0 >= y1 for y0, y1 in strips):
        raise ValueError("a %d-row frame cannot be cut into %d strips of whole 16-row tiles" % (height, world_size))
    return strip_height, strips


def exchange_ipc_handles(local_handle, group=None):
    """all_gather of the per-rank 64-byte CUDA IPC handles (works on the gloo and the nccl backend)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    mine = torch.tensor(list(local_handle), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    return [bytes(t.cpu().tolist()) for t in gathered]


class StripDenoiser(object):
    """The strip of one rank: instance + strip-mode CUDA context.  IN_*/OUT_* strips live in the context's arena."""

    def __init__(self, denoiser, width, height, rank, world_size, device=0, identifier=0, settings=None, halo_rows=None):
        self.denoiser, self.width, self.height, self.identifier = denoiser, width, height, identifier
        self.rank, self.world_size = rank, world_size
        self.device = torch.device("cuda", device)
        self.strip_height, strips = partition_rows(height, world_size)
        self.y0, self.y1 = strips[rank]
        self.instance = nrd.Instance([(identifier, denoiser)])
        self.ctx = nrd.CudaContext(self.instance, width, height, device=device, strip=(self.y0, self.y1), strip_height=self.strip_height,
                                   halo_rows=DEFAULT_HALO_ROWS if halo_rows is None else halo_rows)
        if settings is not None:
            self.instance.set_denoiser_settings(identifier, settings)
        self.names = harness.DENOISER_RESOURCES[denoiser]

    def connect(self, group=None):
        """Multi-process: exchange the IPC handles over torch.distributed and map the peers' arenas."""
        handles = exchange_ipc_handles(self.ctx.ipc_handle(), group)
        self.ctx.connect_peers(self.rank, self.world_size, ipc_handles=handles)

    def connect_local(self, all_strips):
        """Single process (tests): every strip is a context of this process on a device that can address the others."""
        self.ctx.connect_peers(self.rank, self.world_size, arenas=[s.ctx.arena()[0] for s in all_strips])

    def _stream(self, stream):
        return stream.cuda_stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream

    def set_inputs(self, frame, stream=None):
        """frame[name]: full-frame tensors (device or pinned host); only the rows of this strip are copied."""
        s = self._stream(stream)
        for name in self.names:
            if name.startswith("IN_"):
                t = frame[name][self.y0:self.y1]
                self.ctx.copy(getattr(nrd.ResourceType, name), 0, t.data_ptr(), t.stride(0) * t.element_size(), True, s)

    def set_input_strips(self, strips, stream=None):
        """strips[name]: tensors that hold exactly the rows of this strip."""
        s = self._stream(stream)
        for name, t in strips.items():
            self.ctx.copy(getattr(nrd.ResourceType, name), 0, t.data_ptr(), t.stride(0) * t.element_size(), True, s)

    def denoise(self, common_settings, stream=None):
        self.instance.set_common_settings(common_settings)
        return self.ctx.denoise([self.identifier], stream=self._stream(stream))

    def dispatches(self, common_settings):
        self.instance.set_common_settings(common_settings)
        r, raw, n = self.instance.get_compute_dispatches_raw([self.identifier])
        if r != nrd.Result.SUCCESS:
            raise nrd.NrdError("GetComputeDispatches", r)
        return raw, n

    def read_outputs(self, out=None, stream=None):
        """Copies the OUT_* strips into torch tensors (device by default, or the given pinned-host tensors)."""
        s = self._stream(stream)
        res = {}
        for name in self.names:
            if name.startswith("OUT_"):
                fmt, dtype, ch = harness.USER_FORMATS[name]
                rows = self.y1 - self.y0
                t = out[name] if out is not None else torch.empty((rows, self.width, ch) if ch > 1 else (rows, self.width), dtype=dtype, device=self.device)
                self.ctx.copy(getattr(nrd.ResourceType, name), 0, t.data_ptr(), t.stride(0) * t.element_size(), False, s)
                res[name] = t
        return res

    def synchronize(self, stream=None):
        self.ctx.synchronize(self._stream(stream))

    def destroy(self):
        self.ctx.destroy()
        self.instance.destroy()
