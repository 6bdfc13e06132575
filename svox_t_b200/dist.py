"""Multi-GPU data path: rays shard across ranks, the tree and feature table are replicated, and the only exchange
step is the sum of the leaf-feature gradients. One process per GPU, torch.distributed
for the plumbing (NCCL over NVLink/NVSwitch on the B200 box, gloo on CPU for the host-logic tests).

The reference has no multi-GPU support at all; this module is new.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n units (rays, image rows, views) owned by ``rank``."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, rank, world):
    """Slice every field of a Rays namedtuple to this rank's contiguous share."""
    lo, hi = shard_range(rays.origins.shape[0], rank, world)
    return type(rays)(*(t[lo:hi].contiguous() for t in rays))


def shard_image_rows(height, rank, world, tile=8):
    """Row band [y0, y1) of an image for this rank, aligned to the kernels' 8-pixel tiles."""
    lo, hi = shard_range((height + tile - 1) // tile, rank, world)
    return min(lo * tile, height), min(hi * tile, height)


def all_reduce_leaf_grads(grad, group=None, async_op=False):
    """Sum dL/dfeatures[M, D] over the ranks, in place (fp32; NCCL all-reduce on NVLink). No-op for one rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class LeafGradExchange:
    """The leaf-gradient table of one training step, held in SYMMETRIC memory so that the exchange is one hand-written
    kernel (csrc/svoxb_exchange.cu: flag barrier -> in-switch reduction with multimem.ld_reduce / multimem.st over the
    NVSwitch multicast mapping, or peer loads/stores when the fabric offers no multicast -> flag barrier), in place.

        xchg = LeafGradExchange(M, D, device)      # collective: every rank of the group
        grad = xchg.zeroed_table()                 # [M, D] view, zero-filled on the current stream
        ... backward kernels reduce into grad ...
        xchg.all_reduce_()                         # grad now holds the sum over the ranks, on every rank

    torch.distributed._symmetric_memory is the plumbing (allocation + address exchange); the kernel sees raw addresses
    through the C ABI (svoxb_peer_group). Falls back to NCCL's all-reduce on an ordinary tensor when symmetric memory
    is not available (``backend == "nccl"``); a world of one needs neither."""

    FLAG_BYTES = 1 << 16

    def __init__(self, M, D, device, group=None, blocks=None, force_backend=None):
        from . import csrc as _C
        self._C = _C
        self.M, self.D, self.device = int(M), int(D), torch.device(device)
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = self.M * self.D
        self.n_floats = (n + 3) // 4 * 4
        self.backend = "local"
        self._hdl = None
        if self.world > 1:
            self.backend = force_backend or os.environ.get("SVOXB_EXCHANGE", "auto")
            if self.backend in ("auto", "nvls", "p2p"):
                try:
                    self._setup_symmetric(blocks)
                except Exception as e:                              # no symmetric memory on this fabric / build
                    if self.backend != "auto":
                        raise
                    self._why_nccl = f"{type(e).__name__}: {e}"
                    self.backend = "nccl"
        if self._hdl is None:
            self._buf = torch.zeros(self.n_floats, dtype=torch.float32, device=self.device)
        self.table = self._buf[:n].view(self.M, self.D)

    def _setup_symmetric(self, blocks):
        import torch.distributed._symmetric_memory as symm
        lib = self._C.load_library()
        self.blocks = int(blocks or os.environ.get("SVOXB_EXCHANGE_BLOCKS", 0) or lib.svoxb_exchange_max_blocks())
        assert self.blocks * self.world * 4 + 4 <= self.FLAG_BYTES
        total = self.n_floats + self.FLAG_BYTES // 4
        with torch.cuda.device(self.device):
            buf = symm.empty(total, dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(buf, self.group)
            buf.zero_()
            torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                                     # every rank's flags are zero before the first use
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)                # 0: the fabric offers no multicast mapping
        if self.backend == "nvls" and not mc:
            raise RuntimeError("SVOXB_EXCHANGE=nvls but the symmetric allocation has no multicast mapping")
        self._buf, self._hdl = buf, hdl
        self._ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        self._epoch = 1

        def group_for(multicast, blocks):
            return self._C._CPeerGroup(
                rank=self.rank, world=self.world, buffers=ctypes.cast(self._ptrs, ctypes.POINTER(ctypes.c_void_p)),
                multicast=ctypes.c_void_p(multicast), table_offset=0, flags_offset=self.n_floats * 4,
                status_offset=self.n_floats * 4 + self.FLAG_BYTES - 4, blocks=blocks, epoch=1)
        forms = {"p2p": 0}
        if mc:
            forms["nvls"] = mc
        if self.backend in forms:
            self._pg = group_for(forms[self.backend], self.blocks)
        else:
            # auto: both forms move the same bytes through different hardware (in-switch reduction + multicast against
            # plain peer loads / stores); which one wins depends on the number of GPUs, and fewer CTAs than SMs can be
            # faster (8 GPUs: 96 CTAs 0.554 ms, 148 CTAs 0.579 ms) -- time the candidates on this table
            self.tuning = {}
            cands = {}
            for name, mcp in forms.items():
                for blocks in sorted({self.blocks, max(1, self.blocks * 2 // 3)}):
                    cands[f"{name}/{blocks}"] = (name, blocks, group_for(mcp, blocks))
            for key, (name, blocks, pg) in cands.items():
                self._pg = pg
                for _ in range(2):
                    self._launch(None)
                torch.cuda.synchronize(self.device)
                dist.barrier(self.group)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    self._launch(None)
                e1.record()
                torch.cuda.synchronize(self.device)
                t = torch.tensor([e0.elapsed_time(e1) / 4], dtype=torch.float64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)      # every rank sees the same numbers
                self.tuning[key] = float(t.item())
            best = min(self.tuning, key=self.tuning.get)
            self.backend, self.blocks, self._pg = cands[best]
            buf[:self.n_floats].zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)

    def _launch(self, features):
        lib = self._C.load_library()
        self._pg.epoch = self._epoch
        with torch.cuda.device(self.device):
            if features is not None:
                self._C._check(lib.svoxb_exchange_sum_rows(ctypes.byref(self._pg), self.M, self.D,
                                                           self._C._ptr(features), self._C._stream()))
            else:
                self._C._check(lib.svoxb_exchange_sum(ctypes.byref(self._pg), self.n_floats, self._C._stream()))
        self._epoch = (self._epoch + 2) & 0xFFFFFFFF

    def zeroed_table(self):
        """The [M, D] table, zero-filled in stream order (the reference's zeros_like(features), rt_kernel.cu:1415)."""
        self._zeroed_for = None
        self.wait_zero()
        self.table.zero_()
        return self.table

    def zero_async(self, features=None):
        """Zero-fill the table on a side stream, ordered after everything the current stream has queued so far (the last
        reader of the previous step's gradient) and overlapping whatever it queues next -- the forward march, which never
        touches the table and leaves most of the DRAM bandwidth idle. ``table_for_backward`` makes the current stream
        wait for the fill. ``features``: what the coming backward will be for (see ``table_for_backward``)."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._zero_done = torch.cuda.Event()
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            self.table.zero_()
            self._zero_done.record(self._side)
        self._zero_pending = True
        self._zeroed_for = self._C._TensorIdentity(features) if features is not None else None

    def note_zeroed(self, features):
        """The table has just been zero-filled (in stream order) by the per-step table pass of the forward over
        ``features`` (csrc.Activated): the next backward for exactly these features reduces into it as it is."""
        self._zeroed_for = self._C._TensorIdentity(features)

    def wait_zero(self):
        """Make the current stream wait for a pending ``zero_async``."""
        if getattr(self, "_zero_pending", False):
            torch.cuda.current_stream(self.device).wait_event(self._zero_done)
            self._zero_pending = False

    def table_for_backward(self, features):
        """The zero-filled table a backward reduces into: as left by the forward's table pass / ``zero_async`` when that
        zeroed it for these features and nothing has used it since, else zero-filled now. One backward per zero-fill."""
        key, self._zeroed_for = getattr(self, "_zeroed_for", None), None
        self.wait_zero()
        if key is not None and key.matches(features):
            return self.table
        return self.zeroed_table()

    def all_reduce_(self, features=None):
        """Sum the table over the ranks, in place, in stream order on the current stream. Collective.
        ``features``: the table holds the gradient a backward produced for these (replicated) features -- rows with
        sigma <= 0 got no gradient on any rank and are left out of the exchange (svoxb_exchange_sum_rows)."""
        if self.world == 1:
            return self.table
        if self._hdl is None:
            dist.all_reduce(self.table, op=dist.ReduceOp.SUM, group=self.group)
            return self.table
        rows = (features is not None and tuple(features.shape) == (self.M, self.D) and features.is_contiguous()
                and features.dtype == torch.float32 and features.device == self.table.device)
        self._launch(features if rows else None)
        return self.table

    def status(self):
        """0, or 1 + the rank a flag barrier gave up waiting for (synchronises)."""
        if self._hdl is None:
            return 0
        word = self._buf[self.n_floats + self.FLAG_BYTES // 4 - 1:].view(torch.int32)
        return int(word.item())

    def describe(self):
        d = {"backend": self.backend, "world": self.world, "bytes": self.n_floats * 4}
        if self._hdl is not None:
            d["blocks"] = self.blocks
            if getattr(self, "tuning", None):
                d["tuning_ms"] = {k: round(v, 4) for k, v in self.tuning.items()}
        if hasattr(self, "_why_nccl"):
            d["fallback_reason"] = self._why_nccl[:200]
        return d


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value_ms, device):
    """Max of a per-rank scalar (device time in ms) over all ranks."""
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    """Sum of a per-rank scalar over all ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def render_step_sharded(renderer, features, rays, loss_fn, rank, world):
    """One ray-sharded training step: this rank renders its slice, back-propagates, and the leaf gradients are
    summed over ranks. ``loss_fn(out, lo, hi)`` must return this shard's contribution to the global loss.
    Returns (local_loss, features.grad summed over ranks)."""
    local = shard_rays(rays, rank, world)
    lo, hi = shard_range(rays.origins.shape[0], rank, world)
    out = renderer(features, local)
    loss = loss_fn(out, lo, hi)
    loss.backward()
    all_reduce_leaf_grads(features.grad)
    return loss.detach(), features.grad


def render_image_bands(renderer, features, c2w, width, height, fx, fy=None, rank=0, world=1, gather=True):
    """One camera frame sharded over the ranks by horizontal bands (aligned to the kernels' 8-row tiles): every rank
    renders rows [y0, y1) with the tree and features it holds; ``gather=True`` all-gathers the bands so that every rank
    returns the full (height, width, D) image, else this rank's band and its (y0, y1). Inference path (no autograd
    through the gather)."""
    from . import csrc as _C
    y0, y1 = shard_image_rows(height, rank, world)
    D = _C._out_dim(renderer.tree._spec(features, _with_accel=False), renderer._get_options())
    with torch.no_grad():
        if y1 > y0:
            band = renderer.render_persp(features, c2w, width=width, height=height, fx=fx, fy=fy, rows=(y0, y1))
        else:                                          # more ranks than 8-row tiles
            band = features.new_empty((0, width, D))
    if not gather:
        return band, (y0, y1)
    if world == 1 or not dist.is_initialized():
        return band
    bounds = [shard_image_rows(height, r, world) for r in range(world)]
    most = max(b1 - b0 for b0, b1 in bounds)            # all_gather wants equal shapes: pad the bands to the tallest
    padded = features.new_zeros((most, width, D))
    padded[: y1 - y0] = band
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: b1 - b0] for p, (b0, b1) in zip(parts, bounds)], dim=0)


def render_views_sharded(renderer, features, cams, width, height, fx, fy=None, rank=0, world=1):
    """A batch of views split across the ranks (whole views per GPU): returns the images of this
    rank's contiguous share of ``cams`` (list of c2w tensors) and the share's (lo, hi)."""
    lo, hi = shard_range(len(cams), rank, world)
    with torch.no_grad():
        imgs = [renderer.render_persp(features, cams[v], width=width, height=height, fx=fx, fy=fy) for v in range(lo, hi)]
    return imgs, (lo, hi)
