"""Drop-in for ``multi_stft_loss`` (retunegan/models/loss.py:22-62) as a torch.autograd.Function over the
fused forward / backward kernels.

    multi_stft_loss(y, y_g, ret_loss=False, ret_specs=False)
        -> loss | (stft_r, stft_g) | (loss, (stft_r, stft_g))

``y`` (real audio) is treated as a constant, exactly how the reference uses it (retunegan/train.py:139,165:
only ``y_g_hat`` carries gradients).  The spec stacks are ``[B, 2, F, T']`` views of frame-major buffers and
carry gradients back to ``y_g`` (they feed the STFT discriminator, retunegan/train.py:172).
Under DDP every rank holds its own segments; ``ddp_reduce=True`` averages the *reported* loss over ranks with
one NCCL all-reduce of a scalar (the gradient all-reduce is DDP's own, on the generator parameters).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, core
from .config import RETUNEGAN, SpectralConfig

hp: SpectralConfig = RETUNEGAN


def set_hparams(cfg) -> None:
    global hp
    hp = cfg if isinstance(cfg, SpectralConfig) else SpectralConfig.from_hparam(cfg)


_setup_cache = {}


def _setup(cfg: SpectralConfig, dev, B: int, T: int):
    """(plans, handle array, saved bytes, workspace bytes) of one (configuration, device, batch shape): the lookups and
    the two size queries are per-shape constants, not per-step work."""
    key = (cfg, dev.index, B, T)
    hit = _setup_cache.get(key)
    if hit is None:
        lib = _lib.load()
        plans = [core.get_plan(cfg, *p) for p in cfg.multi_stft_params]     # always the Slaney basis (retunegan/audio.py:158)
        handles = (C.c_void_p * len(plans))(*[p.handle for p in plans])
        saved_bytes = lib.sb200_mstft_saved_bytes(handles, len(plans), B, T)
        ws_bytes = lib.sb200_mstft_workspace_bytes(handles, len(plans), B, T)
        if saved_bytes < 0 or ws_bytes < 0:
            _lib.check(-1)
        if len(_setup_cache) > 64:
            _setup_cache.clear()
        hit = _setup_cache[key] = (plans, handles, int(saved_bytes), int(ws_bytes))
    return hit


class PeerLossReducer:
    """Exchange buffers for the in-kernel mean of the loss over the ranks of one box (include/spectral_b200.h, "DDP: the loss
    averaged over the ranks ... INSIDE the reducing kernel").  Building one is a COLLECTIVE over the default process group: every
    rank allocates its buffer, the CUDA IPC handles travel by ``all_gather_object`` and each rank maps the others' buffers
    (NVLink peer access).  ``peers`` (single-process use, tests): the buffers of all ranks as raw device pointers."""

    def __init__(self, rank: int, world: int, peers=None):
        lib = _lib.load()
        self.rank, self.world = int(rank), int(world)
        if not 1 <= self.world <= _lib.MAX_PEERS:
            raise ValueError(f"in-kernel loss reduction needs 1 <= world <= {_lib.MAX_PEERS}")
        self._own, self._opened = None, []
        if peers is not None:
            self.ptrs = [int(p) for p in peers]
            return
        import socket
        import torch.distributed as dist
        own, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(lib.sb200_peer_buffer_create(C.byref(own), handle), "peer_buffer_create")
        self._own = own
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (socket.gethostname(), bytes(handle)))
        ok, self.ptrs = len({h for h, _ in gathered}) == 1, []          # CUDA IPC: one host
        for q, (_, hb) in enumerate(gathered):
            if q == self.rank:
                self.ptrs.append(own.value)
                continue
            p = C.c_void_p()
            if ok and lib.sb200_peer_buffer_open((C.c_ubyte * 64).from_buffer_copy(hb), C.byref(p)) == 0:
                self._opened.append(p)
                self.ptrs.append(p.value)
            else:
                ok = False
                self.ptrs.append(0)
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                       # every rank mapped every buffer, or nobody uses them
        if flag.item() < 1.0:
            self.close()
            raise RuntimeError("peer exchange buffers could not be mapped on every rank")

    @staticmethod
    def create_buffer() -> int:
        """One zeroed exchange buffer in this process (single-process use); free it with ``destroy_buffer``."""
        p = C.c_void_p()
        _lib.check(_lib.load().sb200_peer_buffer_create(C.byref(p), None), "peer_buffer_create")
        return p.value

    @staticmethod
    def destroy_buffer(ptr: int) -> None:
        _lib.check(_lib.load().sb200_peer_buffer_destroy(C.c_void_p(ptr)), "peer_buffer_destroy")

    def descriptor(self, loss_global: torch.Tensor) -> "_lib.PeerReduce":
        d = _lib.PeerReduce()
        for q, p in enumerate(self.ptrs):
            d.peer[q] = p
        d.rank, d.world, d.loss_global = self.rank, self.world, loss_global.data_ptr()
        return d

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.sb200_peer_buffer_close(p)
        self._opened = []
        if self._own is not None:
            lib.sb200_peer_buffer_destroy(self._own)
            self._own = None


_peer_reducers = {}


def ddp_loss_reducer():
    """The process group's PeerLossReducer on the current device, built on first use (collectively: every rank calls
    ``multi_stft_loss(..., ddp_reduce=True)`` the same way), or None when the job is not one NCCL box of <= 8 ranks or the
    mapping failed -- then the loss is reduced with an NCCL all-reduce of a scalar.  ``SB200_DDP_PEER=0`` forces that path."""
    import os
    import torch.distributed as dist
    if os.environ.get("SB200_DDP_PEER", "1") == "0" or not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_backend() != "nccl" or not torch.cuda.is_available():
        return None
    key = torch.cuda.current_device()
    if key not in _peer_reducers:
        red = None
        if 1 < dist.get_world_size() <= _lib.MAX_PEERS:
            try:
                red = PeerLossReducer(dist.get_rank(), dist.get_world_size())
            except Exception:      # not one box, no peer access, IPC unavailable: every rank lands here together (MIN all-reduce)
                red = None
        _peer_reducers[key] = red
    return _peer_reducers[key]


def ddp_reduce_path() -> str:
    """Which way ``ddp_reduce=True`` takes in this process (for logs / bench.py)."""
    return "in-kernel exchange over NVLink peer memory" if ddp_loss_reducer() is not None else "NCCL all-reduce of one scalar"


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


def _as_rows(t: torch.Tensor, dev) -> torch.Tensor:
    """[B, T] / [B, 1, T] (any device / float dtype) -> contiguous float32 [B, T] on `dev`, without touching a tensor that already
    is one (the training step is bound by the host: every avoidable tensor op counts)."""
    if t.dtype is not torch.float32 or t.device != dev or not t.is_contiguous():
        t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
    return t


class _MultiStftFn(torch.autograd.Function):
    """y, y_g: [B, T] or [B, 1, T] (the reference squeezes the depth axis itself, loss.py:27-28; here the two extra views and the
    backward node they add are skipped: the kernels only need the row pointer, B and T)."""

    @staticmethod
    def forward(ctx, y, y_g, cfg: SpectralConfig, want_loss: bool, want_specs: bool, reducer=None):
        lib = _lib.load()
        dev = core.require_cuda()
        ctx.set_materialize_grads(False)   # no zero-filled gradients for outputs nobody differentiated (the real-audio stacks)
        yc = _as_rows(y, dev)
        gc = _as_rows(y_g, dev)
        B, T = gc.shape[0], gc.shape[-1]
        plans, handles, saved_bytes, ws_bytes = _setup(cfg, dev, B, T)
        n_res = len(plans)
        ctx.fused = bool(want_loss and not want_specs and ctx.needs_input_grad[1])
        if ctx.fused:
            # loss-only training step: value and gradient (for a unit upstream gradient) in one pass; backward just scales it
            ws = core._workspace(int(ws_bytes), dev, "mstft")
            loss = torch.empty((), device=dev, dtype=torch.float32)
            grad = torch.empty((B, T), device=dev, dtype=torch.float32)
            if reducer is None:
                _lib.check(lib.sb200_mstft_loss_and_grad(handles, n_res, core.ptr(yc), core.ptr(gc), B, T, core.ptr(loss),
                                                         core.ptr(grad), core.ptr(ws), core.stream_ptr()), "mstft_loss_and_grad")
            else:   # DDP: the value handed out is the mean over the ranks (reduced inside the launch), the gradient the local one
                mean = torch.empty((), device=dev, dtype=torch.float32)
                _lib.check(lib.sb200_mstft_loss_and_grad_ddp(handles, n_res, core.ptr(yc), core.ptr(gc), B, T, core.ptr(loss),
                                                             core.ptr(grad), core.ptr(ws), C.byref(reducer.descriptor(mean)),
                                                             core.stream_ptr()), "mstft_loss_and_grad_ddp")
                loss = mean
            ctx.in_shape, ctx.in_dtype = y_g.shape, y_g.dtype
            ctx.save_for_backward(grad)
            return (loss,)
        saved = torch.empty(int(saved_bytes), device=dev, dtype=torch.uint8)
        ws = core._workspace(int(ws_bytes), dev, "mstft")
        loss = torch.empty((), device=dev, dtype=torch.float32) if want_loss else None
        specs_r = specs_g = ptrs_r = ptrs_g = None
        if want_specs:
            # ONE allocation for the six stacks; each is handed out directly as the [B, 2, F, T'] view (strides (2 T' F, T' F, 1, F))
            # of its frame-major [B, 2, T', F] block: no per-stack allocation, no transpose op per stack and no TransposeBackward
            # node per stack in backward -- on a step that is bound by the host these were a fifth of it
            tfs = [1 + T // p.hop_length for p in plans]
            sizes = [B * 2 * tf * p.F for tf, p in zip(tfs, plans)]
            total = sum(sizes)
            big = torch.empty(2 * total, device=dev, dtype=torch.float32)
            base, offs, o = big.data_ptr(), [], 0
            for n in sizes:
                offs.append(o)
                o += n
            view = lambda off, tf, p: big.as_strided((B, 2, p.F, tf), (2 * tf * p.F, tf * p.F, 1, p.F), off)
            specs_r = [view(off, tf, p) for off, tf, p in zip(offs, tfs, plans)]
            specs_g = [view(total + off, tf, p) for off, tf, p in zip(offs, tfs, plans)]
            ptrs_r = (C.c_void_p * n_res)(*[base + 4 * off for off in offs])
            ptrs_g = (C.c_void_p * n_res)(*[base + 4 * (total + off) for off in offs])
            ctx.spec_dims = [(tf, p.F) for tf, p in zip(tfs, plans)]
        phd_phase = int(cfg.phd_input == "phase")
        if reducer is None or not want_loss:
            _lib.check(lib.sb200_mstft_forward(handles, n_res, core.ptr(yc), core.ptr(gc), B, T, phd_phase, core.ptr(loss),
                                               ptrs_r, ptrs_g,
                                               core.ptr(saved), core.ptr(ws), core.stream_ptr()), "mstft_forward")
        else:
            mean = torch.empty((), device=dev, dtype=torch.float32)
            _lib.check(lib.sb200_mstft_forward_ddp(handles, n_res, core.ptr(yc), core.ptr(gc), B, T, phd_phase, core.ptr(loss),
                                                   ptrs_r, ptrs_g,
                                                   core.ptr(saved), core.ptr(ws), C.byref(reducer.descriptor(mean)),
                                                   core.stream_ptr()), "mstft_forward_ddp")
            loss = mean
        ctx.cfg, ctx.plans, ctx.handles, ctx.ws_bytes = cfg, plans, handles, ws_bytes
        ctx.shape = (B, T)
        ctx.want_loss, ctx.want_specs, ctx.phd_phase = want_loss, want_specs, phd_phase
        ctx.in_shape, ctx.in_dtype = y_g.shape, y_g.dtype
        ctx.save_for_backward(gc, saved)
        outs = []
        if want_loss:
            outs.append(loss)
        if want_specs:
            outs += specs_r + specs_g
            ctx.mark_non_differentiable(*specs_r)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.fused:
            (grad,) = ctx.saved_tensors
            g0 = grads[0]
            if g0 is None:
                return None, None, None, None, None, None
            if g0.dtype is not torch.float32 or g0.device != grad.device:
                g0 = g0.to(device=grad.device, dtype=torch.float32)
            g = (g0 * grad).view(ctx.in_shape)
            return None, (g if ctx.in_dtype is torch.float32 else g.to(ctx.in_dtype)), None, None, None, None
        lib = _lib.load()
        gc, saved = ctx.saved_tensors
        B, T = ctx.shape
        n_res = len(ctx.plans)
        dev = gc.device
        i = 0
        g_loss = None
        if ctx.want_loss:
            g_loss = grads[0]
            i = 1
            if g_loss is not None:
                g_loss = _as_rows(g_loss, dev)
        g_specs = None
        if ctx.want_specs:
            gs = grads[i + n_res: i + 2 * n_res]
            if any(g is not None for g in gs):
                # the kernels read frame-major [B, 2, T', F] memory: a gradient that already has it behind its [B, 2, F, T'] shape
                # is used in place, anything else (e.g. a contiguous [B, 2, F, T'] gradient out of a convolution) is copied over
                g_specs = []
                for g, (tf, F) in zip(gs, ctx.spec_dims):
                    if g is not None and not (g.dtype is torch.float32 and g.device == dev and
                                              g.stride() == (2 * tf * F, tf * F, 1, F)):
                        g = g.to(device=dev, dtype=torch.float32).transpose(2, 3).contiguous()
                    g_specs.append(g)
        g_yg = torch.empty((B, T), device=dev, dtype=torch.float32)
        ws = core._workspace(ctx.ws_bytes, dev, "mstft")
        _lib.check(lib.sb200_mstft_backward(ctx.handles, n_res, core.ptr(gc), B, T, ctx.phd_phase, core.ptr(g_loss),
                                            _ptr_array(g_specs) if g_specs is not None else None, core.ptr(saved),
                                            core.ptr(g_yg), core.ptr(ws), core.stream_ptr()), "mstft_backward")
        g_yg = g_yg.view(ctx.in_shape)
        return None, (g_yg if ctx.in_dtype is torch.float32 else g_yg.to(ctx.in_dtype)), None, None, None, None


class _GlobalMean(torch.autograd.Function):
    """Value = mean of the rank-local losses (one NCCL all-reduce of a scalar), gradient = the local one: what DDP training logs
    and back-propagates (retunegan/train.py averages nothing itself; DDP averages the generator gradients later)."""

    @staticmethod
    def forward(ctx, loss):
        red = loss.detach() / torch.distributed.get_world_size()
        torch.distributed.all_reduce(red, op=torch.distributed.ReduceOp.SUM)
        return red

    @staticmethod
    def backward(ctx, g):
        return g


def multi_stft_loss(y, y_g, ret_loss=False, ret_specs=False, ddp_reduce=False):
    """retunegan/models/loss.py:22-62 (same arguments and return structure)."""
    if not (ret_loss or ret_specs):
        raise RuntimeError("multi_stft_loss: neither ret_loss nor ret_specs")   # bare `raise` at loss.py:62
    if hp.phd_input not in ("stft", "phase"):
        raise RuntimeError(f"unknown phd_input {hp.phd_input!r}")               # bare `raise` at loss.py:48
    if y.shape != y_g.shape or not (y.dim() == 2 or (y.dim() == 3 and y.shape[1] == 1)):   # [B, 1, T] is taken as [B, T]
        raise ValueError(f"expected matching [B, T] / [B, 1, T] inputs, got {tuple(y.shape)} and {tuple(y_g.shape)}")
    ddp = bool(ddp_reduce and ret_loss and torch.distributed.is_available() and torch.distributed.is_initialized())
    reducer = ddp_loss_reducer() if ddp else None
    outs = _MultiStftFn.apply(y, y_g, hp, bool(ret_loss), bool(ret_specs), reducer)
    n_res = len(hp.multi_stft_params)
    i = 0
    loss = None
    if ret_loss:
        loss = outs[0]
        i = 1
        if ddp and reducer is None:
            loss = _GlobalMean.apply(loss)   # value = global mean, grad = local
    if ret_specs:
        stft_r = list(outs[i:i + n_res])          # [B, 2, F, T'] views of frame-major buffers (loss.py:44-45 stacks)
        stft_g = list(outs[i + n_res:i + 2 * n_res])
    if ret_loss and ret_specs:
        return loss, (stft_r, stft_g)
    elif ret_loss:
        return loss
    return (stft_r, stft_g)


class _PoolLossFn(torch.autograd.Function):
    """Value and d/dy_g (for a unit upstream gradient) of a waveform max-pool loss in one launch; backward scales it."""

    @staticmethod
    def forward(ctx, y, y_g, mode: int, pool_k: int):
        lib = _lib.load()
        dev = core.require_cuda()
        yc = y.detach().to(device=dev, dtype=torch.float32).contiguous()
        gc = y_g.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, T = gc.shape
        need_grad = ctx.needs_input_grad[1]
        loss = torch.empty((), device=dev, dtype=torch.float32)
        grad = torch.empty((B, T), device=dev, dtype=torch.float32) if need_grad else None
        ws = core._workspace(int(lib.sb200_pool_loss_workspace_bytes()), dev, "pool_loss")
        _lib.check(lib.sb200_pool_loss(core.ptr(yc), core.ptr(gc), B, T, int(pool_k), int(mode), core.ptr(loss), core.ptr(grad),
                                       core.ptr(ws), core.stream_ptr()), "pool_loss")
        ctx.in_shape, ctx.in_dtype = y_g.shape, y_g.dtype
        if need_grad:
            ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (grad,) = ctx.saved_tensors
        return None, (grad * g_loss).to(ctx.in_dtype).reshape(ctx.in_shape), None, None


def _pool_loss(y, y_g, mode):
    if y.shape != y_g.shape or y.dim() not in (2, 3) or (y.dim() == 3 and y.shape[1] != 1):
        raise ValueError(f"expected matching [B, T] / [B, 1, T] inputs, got {tuple(y.shape)} and {tuple(y_g.shape)}")
    if y.shape[-1] < hp.envelope_pool_k:
        raise RuntimeError("max_pool1d: output size is too small")           # what nn.MaxPool1d raises
    return _PoolLossFn.apply(y.reshape(y.shape[0], -1), y_g.reshape(y_g.shape[0], -1), mode, hp.envelope_pool_k)


def envelope_loss(y, y_g):
    """retunegan/models/loss.py:66-72: ``mean|MaxPool(y) - MaxPool(y_g)| + mean|MaxPool(-y) - MaxPool(-y_g)|``."""
    return _pool_loss(y, y_g, 0)


def dynamic_loss(y, y_g):
    """retunegan/models/loss.py:76-82: ``mean||MaxPool(y) + MaxPool(-y)| - |MaxPool(y_g) + MaxPool(-y_g)||``."""
    return _pool_loss(y, y_g, 1)
