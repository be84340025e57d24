"""CUDA-graph execution of the decode path.

The eager path issues ~870 kernel launches per 768x768 image from Python; on a B200 that is launch-bound.  The decode
has a fixed structure per (batch, H, W), broken only by the four host rANS calls, so it is captured once into five
graphs (one graph for the device-resident leg) and replayed:

    G0 : FSQ codes -> hyper-synthesis -> prior fusion -> reduction -> index kernel(step 0) -> D2H indices
    Gk : H2D symbols(k-1) -> dequant(k-1) -> adaptor_k + prior net -> index kernel(k) -> D2H indices      k = 1..3
    Gs : (side stream, overlapping the loop and the host rANS) semantic adaptor -> sem_up half of g_s
    G4 : H2D symbols(3) -> dequant(3) -> main half of g_s -> UNet -> x0 -> VAE -> D2H image

Host<->device copies use fixed pinned buffers and are graph nodes; TMA descriptors are encoded at capture time against
the (static) addresses of the graph's private memory pool.
"""
import time

import numpy as np
import torch

from . import bitstream, ops
from .entropy_models import StreamDecoder


class GraphedDecoder:
    def __init__(self, model, batch, height, width, lane_base=0):
        """lane_base: first of the two scratch lanes (split-K workspace, GroupNorm accumulators) this decoder's graphs
        are captured against; decoders that replay CONCURRENTLY (PipelinedDecoder) must not share lanes."""
        assert height % 64 == 0 and width % 64 == 0, "graphs are keyed by the padded size"
        self.model, self.codec = model, model.codec_model
        self.B, self.H, self.W = batch, height, width
        self.lane_base = lane_base
        dev = model.device
        self.dev = dev
        self.hz, self.wz = height // 64, width // 64
        self.h16, self.w16 = height // 16, width // 16
        self.nsym = 32 * self.h16 * self.w16
        B = batch
        self.z_host = torch.zeros((B, self.hz, self.wz), dtype=torch.int32, pin_memory=True)
        self.z_dev = torch.zeros((B, self.hz, self.wz), dtype=torch.int32, device=dev)
        self.idx_host = torch.zeros((B, self.nsym), dtype=torch.int16, pin_memory=True)
        self.sym_host = torch.zeros((B, self.nsym), dtype=torch.int16, pin_memory=True)
        self.idx_dev = torch.zeros((B, 32, self.h16, self.w16), dtype=torch.int16, device=dev)
        self.sym_dev = torch.zeros((B, 32, self.h16, self.w16), dtype=torch.int16, device=dev)
        self.img_host = torch.zeros((B, 3, height, width), dtype=torch.float32, pin_memory=True)
        self.syms_res = [torch.zeros((B, 32, self.h16, self.w16), dtype=torch.int16, device=dev) for _ in range(4)]
        self.side = torch.cuda.Stream(device=dev)
        self.sem_done = torch.cuda.Event()
        self.last_rans_ms = 0.0
        self.graphs = None
        self.g_res = None
        self.img_dev = None
        self._keep = []

    # ---------------------------------------------------------------------------------------------
    def _seg0(self):
        c = self.codec
        self.common, self.z_sem = c.hyper(self.z_dev)
        self.params = c.prior.init_params(self.common)
        self.sm = [self.common, None, None, None]
        ops.scale_to_index(self.common[..., :128], c._lut(), 0, self.idx_dev)
        self.idx_host.copy_(self.idx_dev.view(self.B, self.nsym), non_blocking=True)

    def _segk(self, k):
        c = self.codec
        self.sym_dev.view(self.B, self.nsym).copy_(self.sym_host, non_blocking=True)
        ops.dequant_accum(self.sym_dev, self.sm[k - 1][..., 128:], self.params[..., :128], k - 1)
        self.sm[k] = c.prior.step(k, self.params)
        ops.scale_to_index(self.sm[k][..., :128], c._lut(), k, self.idx_dev)
        self.idx_host.copy_(self.idx_dev.view(self.B, self.nsym), non_blocking=True)

    def _seg_sem(self):
        """Semantic branch (needs only z): semantic adaptor + the sem_up half of g_s.  Replayed on a side stream so
        it fills the GPU while the host runs rANS; uses its own scratch lane."""
        c = self.codec
        ops.SCRATCH_LANE = self.lane_base + 1
        try:
            codes = ops.fsq_codes(self.z_dev)
            z_sem = ops.igemm(codes, c.hyper.feat_in, act=ops.ACT_LRELU, slope=0.01)
            self.y_sem = c.semantic_adaptor(z_sem)
            self.cat = c.dec.alloc_cat(self.B, self.h16, self.w16, self.dev)
            c.dec.sem_path(self.y_sem, self.cat)
            # the UNet's 16 cross-attention K | V projections depend on the hyperprior tokens only
            b, hz, wz, cs = self.y_sem.shape
            self.ctx_kv = self.model.feedforward_model.project_ctx(self.y_sem.reshape(b, hz * wz, cs))
        finally:
            ops.SCRATCH_LANE = self.lane_base

    def _seg4(self):
        c = self.codec
        self.sym_dev.view(self.B, self.nsym).copy_(self.sym_host, non_blocking=True)
        ops.dequant_accum(self.sym_dev, self.sm[3][..., 128:], self.params[..., :128], 3)
        x_hat = c.dec.main_path(self.params[..., :128], self.cat)
        self.img_dev = self.model.generate(x_hat, self.y_sem, ctx_kv=self.ctx_kv)
        self.img_host.copy_(self.img_dev, non_blocking=True)

    def _segments(self):
        return [self._seg0, lambda: self._segk(1), lambda: self._segk(2), lambda: self._segk(3), self._seg_sem,
                self._seg4]

    def capture(self):
        """Warm up eagerly once (function attributes, allocator), then capture the five segments into one pool."""
        if self.graphs is not None:
            return
        ops.SCRATCH_LANE = self.lane_base
        try:
            self._capture()
        finally:
            ops.SCRATCH_LANE = 0

    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        traces = []
        with torch.cuda.stream(s), torch.no_grad():
            for seg in self._segments():
                ops.WEIGHT_TRACE = []
                try:
                    seg()
                finally:
                    traces.append(ops.WEIGHT_TRACE)
                    ops.WEIGHT_TRACE = None
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # the last launch of a main-path segment hints the first weights of the next one (the four host rANS calls
        # sit in between: the prefetch runs under them); the semantic segment (index 4) is its own chain
        order = [0, 1, 2, 3, 5]
        for a, b in zip(order[:-1], order[1:]):
            nxt = next((w for w in traces[b] if w[1] > 0), None)
            if nxt is not None:
                traces[a] = traces[a] + [nxt]
        from . import lib
        pool = torch.cuda.graph_pool_handle()
        self.graphs = []
        n0 = lib.launch_count()
        sem_pool = torch.cuda.graph_pool_handle()     # the semantic graph runs CONCURRENTLY with G1..G3: its memory
        with torch.no_grad():                          # must never alias their (recycled) temporaries
            for i, seg in enumerate(self._segments()):
                g = torch.cuda.CUDAGraph()
                ops.WEIGHT_PLAY = [traces[i], 0]
                try:
                    with torch.cuda.graph(g, pool=sem_pool if i == 4 else pool):
                        seg()
                finally:
                    ops.WEIGHT_PLAY = None
                self.graphs.append(g)
        self.launches = lib.launch_count() - n0
        torch.cuda.synchronize()

    def capture_resident(self):
        if self.g_res is not None:
            return
        from . import lib
        assert self.lane_base == 0, "the resident leg is only captured on the default lanes"
        with torch.no_grad():
            ops.WEIGHT_TRACE = []
            try:
                self.model.decode_resident(self.z_dev, self.syms_res)
            finally:
                trace, ops.WEIGHT_TRACE = ops.WEIGHT_TRACE, None
            torch.cuda.synchronize()
            n0 = lib.launch_count()
            self.g_res = torch.cuda.CUDAGraph()
            ops.WEIGHT_PLAY = [trace, 0]
            try:
                with torch.cuda.graph(self.g_res):
                    self.img_res = self.model.decode_resident(self.z_dev, self.syms_res)
            finally:
                ops.WEIGHT_PLAY = None
            self.launches_res = lib.launch_count() - n0      # kernels of this library inside one replay
        torch.cuda.synchronize()

    def capture_z_only(self):
        """z-only model: indices -> image is one graph (no y stream, no host round trip), D2H of the image included."""
        if getattr(self, "g_z", None) is not None:
            return
        from . import lib
        assert self.lane_base == 0
        with torch.no_grad():
            x_hat, y_sem = self.codec.decode_z_only(self.z_dev)          # eager warm-up
            self.model.generate(x_hat, y_sem)
            torch.cuda.synchronize()
            n0 = lib.launch_count()
            self.g_z = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_z):
                x_hat, y_sem = self.codec.decode_z_only(self.z_dev)
                self.img_z = self.model.generate(x_hat, y_sem)
                self.img_host.copy_(self.img_z, non_blocking=True)
            self.launches_z = lib.launch_count() - n0
        torch.cuda.synchronize()

    def run_z_only(self):
        self.capture_z_only()
        self.g_z.replay()
        return self.img_z

    def decode_z_only(self, z_idx):
        """z_idx int [B,hz,wz] (host or device) -> fresh device image; the pinned host copy is in `img_host`."""
        self.capture_z_only()
        if z_idx.is_cuda:
            self.z_dev.copy_(z_idx.to(torch.int32), non_blocking=True)
        else:
            self.z_host.copy_(z_idx.to(torch.int32))
            self.z_dev.copy_(self.z_host, non_blocking=True)
        self.g_z.replay()
        torch.cuda.current_stream().synchronize()
        return self.img_z.clone()

    # ---------------------------------------------------------------------------------------------
    def decode(self, streams):
        """streams: B reference-format containers of this padded size -> pinned host images [B,3,H,W] fp32
        (padded; crop with the returned headers).  Host rANS runs between graph replays."""
        assert len(streams) == self.B
        self.capture()
        c = self.codec
        hdrs = [bitstream.decode_i(s, c.index_unit_length, c.ds) for s in streams]
        for i, d in enumerate(hdrs):
            assert (d["pad_height"], d["pad_width"]) == (self.H, self.W)
            self.z_host[i] = torch.from_numpy(bitstream.unpack_indices(d["bit_stream_z"], self.hz * self.wz,
                                                                       c.index_unit_length).reshape(self.hz, self.wz))
        decoders = [StreamDecoder(c.entropy_coder, d["bit_stream_y"], c.gaussian_encoder.cdf_group_index) for d in hdrs]
        stream = torch.cuda.current_stream()
        self.z_dev.copy_(self.z_host, non_blocking=True)
        # semantic branch on the side stream: overlaps the prior loop and the host rANS calls
        self.side.wait_stream(stream)
        with torch.cuda.stream(self.side):
            self.graphs[4].replay()
            self.sem_done.record(self.side)
        ip, sp, n = self.idx_host.data_ptr(), self.sym_host.data_ptr(), self.nsym
        t_rans = 0.0
        for k in range(4):
            self.graphs[k].replay()
            stream.synchronize()
            t0 = time.perf_counter()
            if self.B == 1:
                decoders[0].decode_into(ip, n, sp)
            else:
                list(c._pool.map(lambda i: decoders[i].decode_into(ip + 2 * i * n, n, sp + 2 * i * n), range(self.B)))
            t_rans += time.perf_counter() - t0
        stream.wait_event(self.sem_done)
        self.graphs[5].replay()
        stream.synchronize()
        self.last_rans_ms = t_rans * 1e3
        return self.img_host, hdrs

    def release(self):
        """Drops the captured graphs, their private memory pools and the pinned buffers (LRU eviction, weight reload)."""
        torch.cuda.synchronize(self.dev)
        self.graphs = self.g_res = self.g_z = self.img_dev = self.img_res = self.img_z = None
        for name in ("common", "z_sem", "params", "sm", "y_sem", "cat", "ctx_kv", "z_host", "z_dev", "idx_host", "sym_host",
                     "idx_dev", "sym_dev", "img_host", "syms_res"):
            self.__dict__.pop(name, None)
        self._keep = []

    def set_resident_inputs(self, z_idx, syms):
        self.z_dev.copy_(z_idx)
        for a, b in zip(self.syms_res, syms):
            a.copy_(b)

    def run_resident(self):
        self.capture_resident()
        self.g_res.replay()
        return self.img_res


class PipelinedDecoder:
    """Several images in flight on one GPU.  One image is a serial chain -- four
    (graph replay -> D2H indices -> host rANS -> H2D symbols) rounds and the final graph -- during which the GPU idles
    while the host decodes and the host idles while the GPU runs, and ~250 of its launches are too small to fill the
    GPU.  `depth` GraphedDecoder slots, each with its own CUDA stream, pinned buffers, graphs and scratch lanes, are
    driven by `depth` host threads that pull streams from a queue: the rANS call releases the GIL, so one slot's
    entropy decode overlaps another slot's kernels, and small kernels of different images share the GPU.  Results are
    bit-identical to decoding the images one by one (same graphs, same kernels)."""

    def __init__(self, model, height, width, depth=2):
        self.model, self.H, self.W, self.depth = model, height, width, depth
        self.slots = [GraphedDecoder(model, 1, height, width, lane_base=2 * (i + 1)) for i in range(depth)]
        self.streams = [torch.cuda.Stream(device=model.device) for _ in range(depth)]
        for sl, st in zip(self.slots, self.streams):          # capture up front, one after the other
            with torch.cuda.stream(st):
                sl.capture()
        torch.cuda.synchronize()

    def release(self):
        for sl in self.slots:
            sl.release()
        self.slots = []

    def decode_many(self, streams, out=None):
        """streams: reference-format containers of this padded size -> list of fp32 [1,3,H,W] host tensors (cropped),
        in input order.  `out`: optional preallocated pinned tensor [len(streams),3,H,W] to receive the padded images."""
        import queue
        import threading
        n = len(streams)
        q = queue.SimpleQueue()
        for i in range(n):
            q.put(i)
        results = [None] * n
        errors = []

        def worker(k):
            slot, st = self.slots[k], self.streams[k]
            try:
                with torch.cuda.stream(st), torch.no_grad():
                    while True:
                        try:
                            i = q.get_nowait()
                        except queue.Empty:
                            return
                        host, hdrs = slot.decode([streams[i]])
                        d = hdrs[0]
                        if out is not None:
                            out[i].copy_(host[0])
                            results[i] = out[i:i + 1, :, :d["height"], :d["width"]]
                        else:
                            results[i] = host[:, :, :d["height"], :d["width"]].clone()
            except Exception as e:                             # surfaced to the caller below
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(self.depth)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results
