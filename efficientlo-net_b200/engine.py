"""CUDA-graph engine: the whole PWCLO forward (≈40 kernel launches) replayed as one graph launch.

Problem sizes on this path are tiny (3600 / 904 / 228 / 116 queries per level), so at small batch the
forward is launch- and latency-bound; capturing it removes the per-launch host cost.  Inputs are
copied into static device buffers (from pinned host memory when they arrive on the host) and the
outputs are static tensors that the next replay overwrites.
"""
import torch

from . import pwclo_model
from .params import init_params, make_perms
from .store import ParamStore


def _order_after_producer(stream, device, *tensors):
    """Device-resident inputs were produced on the caller's current stream: the copy on `stream` must wait for it,
    and the caching allocator must not recycle them while that copy is in flight."""
    cuda = [t for t in tensors if t is not None and t.is_cuda]
    if not cuda:
        return
    stream.wait_stream(torch.cuda.current_stream(device))
    for t in cuda:
        t.record_stream(stream)


class PWCLOEngine:
    def __init__(self, batch_size, H_input=64, W_input=1800, num_points=150000, params=None, perms=None,
                 device="cuda:0", use_graph=True, packed=False, band=None):
        """packed=False: the static input buffer has the reference's placeholder shape (B, 2N, 6) (pwclo_model.py:19).
        packed=True: it is (B, 2N, 3) -- xyz only, the 12 of 24 bytes per point the path reads -- and load() also
        accepts the two frames as (B, n1, 3) / (B, n2, 3) prefixes without their zero padding (the padding is done on
        the device: rows past the prefix stay, or are reset to, zero)."""
        self.B, self.H, self.W, self.N = batch_size, H_input, W_input, num_points
        self.band = band              # pwclo_model.RowBand: this engine computes one row band of every pair (B = 1);
                                      # all ranks of the band's group must load the same pair and call run() together
        self.packed = bool(packed)
        self.filled = [0, 0]          # rows of each frame's slab that may hold points (packed prefix uploads)
        self.device = torch.device(device)
        self.store = params if isinstance(params, ParamStore) else ParamStore(
            params if params is not None else init_params(0), self.device)
        self.perms = perms if perms is not None else make_perms(0)
        self.perms = {k: v.to(self.device) for k, v in self.perms.items()}
        self.pc = torch.zeros(batch_size, 2 * num_points, 3 if self.packed else 6, device=self.device)
        self.T_gt = torch.eye(4, device=self.device).expand(batch_size, 4, 4).contiguous()
        self.use_graph = use_graph
        self.graph = None
        self.outputs = None
        self.stream = torch.cuda.Stream(self.device)
        self.launches_per_forward = None

    def _forward(self):
        # own scratch name space: forwards of different engines may overlap on different streams
        prev, self.store.scratch_ns = self.store.scratch_ns, id(self)
        try:
            return pwclo_model.get_model(self.pc, self.H, self.W, self.T_gt, None, None, False, params=self.store,
                                         perms=self.perms, band=self.band)
        finally:
            self.store.scratch_ns = prev

    def capture(self):
        """Warm up eagerly (allocates scratch, sets kernel attributes), then capture one forward."""
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(2 if self.use_graph else 3):
                    self.outputs = self._forward()
            self.stream.synchronize()
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    self.outputs = self._forward()
            torch.cuda.synchronize(self.device)
        return self

    def load(self, point_cloud, T_gt=None, non_blocking=True, stream=None):
        """Stage one batch into the static device buffer, on the engine's stream (or `stream`).
        point_cloud: a (B, 2N, 6) tensor (or (B, 2N, 3) for a packed engine), host (ideally pinned) or device; or, for
        a packed engine, a pair (xyz_f1 (B, n1, 3), xyz_f2 (B, n2, 3)) with n1, n2 <= N: only the rows that hold
        points travel, the zero padding of main.py:327-333 / kitti_dataset.py:76-80 happens here on the device."""
        stream = stream if stream is not None else self.stream
        frames = point_cloud if isinstance(point_cloud, (tuple, list)) else None
        _order_after_producer(stream, self.device, *(frames if frames is not None else (point_cloud,)), T_gt)
        with torch.cuda.stream(stream):
            if frames is None:
                if point_cloud.shape[-1] != self.pc.shape[-1]:
                    raise ValueError("this engine takes (B, 2N, %d) clouds%s" % (
                        self.pc.shape[-1], " or a pair of (B, n, 3) frame prefixes" if self.packed else
                        "; build it with packed=True for xyz-only uploads"))
                self.pc.copy_(point_cloud, non_blocking=non_blocking)
                self.filled = [self.N, self.N]
            else:
                if not self.packed:
                    raise ValueError("frame prefixes need an engine built with packed=True")
                for f, xyz in enumerate(frames):
                    n = xyz.shape[1]
                    if xyz.shape[0] != self.B or xyz.shape[2] != 3 or n > self.N:
                        raise ValueError("frame %d: expected (%d, n <= %d, 3), got %s" % (f + 1, self.B, self.N, tuple(xyz.shape)))
                    slab = self.pc[:, f * self.N:(f + 1) * self.N]
                    slab[:, :n].copy_(xyz, non_blocking=non_blocking)
                    if n < self.filled[f]:
                        slab[:, n:self.filled[f]].zero_()          # rows a longer earlier frame left behind
                    self.filled[f] = n
            if T_gt is not None:
                self.T_gt.copy_(T_gt, non_blocking=non_blocking)

    def run(self):
        """One forward on the engine's stream (asynchronous).  Returns the static 11-tuple of outputs."""
        if self.graph is None and self.use_graph:
            self.capture()
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.outputs = self._forward()
        return self.outputs

    def infer(self, point_cloud, T_gt=None):
        """The call a user makes: host or device batch in, (q (B,4), t (B,3)) of the finest level out on
        the host (synchronises)."""
        self.load(point_cloud, T_gt)
        out = self.run()
        with torch.cuda.stream(self.stream):
            q, t = out[0].to("cpu", non_blocking=True), out[1].to("cpu", non_blocking=True)
        self.stream.synchronize()
        return q, t


class PWCLOPipeline:
    """Streaming inference from host memory: while captured forwards run, the next batches are uploaded
    on a copy stream (`depth` engine instances = `depth` input buffers + graphs, used round-robin).  Every
    batch still pays its own host->device copy and device->host read of (q, t); they just overlap the compute
    of the neighbouring batches, as a deployment that consumes a LiDAR stream would run it.
    `streams` > 1 keeps that many forwards in flight at once on separate compute streams: one forward of a
    single frame pair is a chain of ~40 dependent single-wave kernels that leaves most of the 148 SMs idle,
    so independent frame pairs overlap almost freely (results are still delivered in order)."""

    def __init__(self, batch_size, H_input=64, W_input=1800, num_points=150000, params=None, perms=None,
                 device="cuda:0", depth=None, streams=1, packed=False):
        depth = depth if depth is not None else 2 * max(1, streams)
        self.device = torch.device(device)
        store = params if isinstance(params, ParamStore) else ParamStore(
            params if params is not None else init_params(0), self.device)
        # several forwards in flight: full 128-row tiles (least SM time per forward) instead of spreading every
        # small call over all SMs (shortest single forward); the policy is read when the graphs are captured
        from . import _lib
        prev = _lib.lib().elo_get_tile_policy()
        if streams > 1:
            _lib.set_tile_policy(1)
        try:
            self.engines = [PWCLOEngine(batch_size, H_input, W_input, num_points, params=store, perms=perms,
                                        device=device, packed=packed).capture() for _ in range(depth)]
        finally:
            _lib.set_tile_policy(prev)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_streams = [torch.cuda.Stream(self.device) for _ in range(max(1, streams))]
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        self.results = [(torch.empty(batch_size, 4).pin_memory(), torch.empty(batch_size, 3).pin_memory())
                        for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        for i, e in enumerate(self.consumed):
            e.record(self.compute_streams[i % len(self.compute_streams)])

    def _upload(self, slot, pc, T_gt):
        eng = self.engines[slot]
        self.copy_stream.wait_event(self.consumed[slot])          # the previous forward on this slot has read its input
        eng.load(pc, T_gt, non_blocking=True, stream=self.copy_stream)
        self.uploaded[slot].record(self.copy_stream)

    def _compute(self, slot):
        eng = self.engines[slot]
        cs = self.compute_streams[slot % len(self.compute_streams)]
        with torch.cuda.stream(cs):
            cs.wait_event(self.uploaded[slot])
            eng.graph.replay()
            self.consumed[slot].record(cs)
            q, t = self.results[slot]
            q.copy_(eng.outputs[0], non_blocking=True)
            t.copy_(eng.outputs[1], non_blocking=True)
            self.done[slot].record(cs)

    def run(self, batches):
        """batches: iterable of (point_cloud, T_gt or None); point_cloud as PWCLOEngine.load takes it -- a (B,2N,6)
        pinned host tensor, or for a packed pipeline (B,2N,3) / a pair of (B,n,3) frame prefixes.  Yields (q, t) host
        tensors per batch, in order (each valid until `depth` more batches have been consumed)."""
        depth = len(self.engines)
        pending = []
        it = iter(batches)
        i = 0
        for pc, T in it:
            slot = i % depth
            if len(pending) == depth:
                s = pending.pop(0)
                self.done[s].synchronize()
                yield self.results[s]
            self._upload(slot, pc, T)
            self._compute(slot)
            pending.append(slot)
            i += 1
        for s in pending:
            self.done[s].synchronize()
            yield self.results[s]
