"""Host-side plumbing shared by the model hosts: tensor intake (DLPack / numpy / torch), arenas, the Keras-style
Adam front end over ts_optim, and the MultiWorkerMirroredStrategy shim over torch.distributed (NCCL).

PyTorch is used for device memory, streams and the process group only; all math runs in libtethys.so.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

_DT = {_lib.TS_F32: torch.float32, _lib.TS_BF16: torch.bfloat16, _lib.TS_I32: torch.int32, _lib.TS_I64: torch.int64}


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_device(x, dtype, device):
    """Accept torch / numpy / any DLPack producer (e.g. a TensorFlow tensor via tf.experimental.dlpack) and return a
    contiguous CUDA torch tensor of `dtype` on `device` (zero-copy when it already is one)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(np.asarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def view_from_ptr(p, shape, ts_dtype, device):
    """Wrap library-owned workspace memory as a torch tensor (no copy) through the CUDA array interface."""
    dt = _DT[ts_dtype]
    n = int(np.prod(shape))
    typestr = {torch.float32: "<f4", torch.bfloat16: "<i2", torch.int32: "<i4", torch.int64: "<i8"}[dt]

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(p), False), "version": 3, "strides": None}
    t = torch.as_tensor(h, device=device)
    if dt == torch.bfloat16:
        t = t.view(torch.bfloat16)
    return t.view(*[int(s) for s in shape])


class ProgramBase:
    """Arenas (torch memory) + a native model-program handle (ts_w2v / ts_whisper). `prefix` selects the C-ABI family."""

    def __init__(self, prefix, precision, device, create):
        self.prefix = prefix
        self.ctx = _lib.context(device)
        self.lib = self.ctx.lib
        self.device = torch.device("cuda", device)
        self.precision = {"fp32": _lib.TS_F32, "float32": _lib.TS_F32, "bf16": _lib.TS_BF16, "bfloat16": _lib.TS_BF16}[precision]
        h = C.c_void_p()
        self.ctx.check(create(self.ctx.h, self.precision, C.byref(h)))
        self.h = h
        self.n = int(self._f("arena_elems")(h))
        self.params = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.params_lp = torch.zeros(self.n, dtype=torch.bfloat16, device=self.device) if self.precision == _lib.TS_BF16 else None
        self.workspace = None
        self.ws_key = None
        self.weights_synced = False
        self.info = {}
        name = C.create_string_buffer(256)
        off, nd, ld = C.c_int64(), C.c_int32(), C.c_int64()
        shape = (C.c_int64 * 4)()
        for i in range(self._f("num_params")(h)):
            self.ctx.check(self._f("param_info")(h, i, name, 256, C.byref(off), C.byref(nd), shape, C.byref(ld)))
            shp = tuple(int(shape[j]) for j in range(nd.value))
            self.info[name.value.decode()] = (int(off.value), shp, int(ld.value))
        self.stage_ends = [int(self._f("stage_end")(h, s)) for s in range(self._f("num_stages")(h))]
        self._optim = None

    def _f(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def view(self, arena, name):
        off, shp, ld = self.info[name]
        numel = int(np.prod(shp))
        if len(shp) == 2 and ld != numel:     # column slice of a fused / padded block (q/k/v, lm_head)
            return arena.as_strided(shp, (ld, 1), off)
        return arena[off:off + numel].view(*shp)

    def make_optim(self):
        if self._optim is None:
            names = list(self.info)
            n = len(names)
            offs = (C.c_int64 * n)(); rows = (C.c_int32 * n)(); cols = (C.c_int32 * n)(); lds = (C.c_int64 * n)()
            for i, k in enumerate(names):
                off, shp, ld = self.info[k]
                numel = int(np.prod(shp))
                r = shp[0] if (len(shp) == 2 and ld != numel) else 1
                offs[i], rows[i], cols[i], lds[i] = off, r, numel // r, ld if r > 1 else numel
            o = C.c_void_p()
            self.ctx.check(self.lib.ts_optim_create(self.ctx.h, n, offs, rows, cols, lds, self.n, C.byref(o)))
            self._optim = o
        return self._optim

    def make_optim_range(self, a0, a1):
        """ts_optim over the variables that live in the arena range [a0, a1) (one all-reduce bucket); ranges must be cut at
        backward-stage ends so that fused blocks (q/k/v) are never split."""
        names = [k for k in self.info if a0 <= self.info[k][0] < a1]
        n = len(names)
        offs = (C.c_int64 * n)(); rows = (C.c_int32 * n)(); cols = (C.c_int32 * n)(); lds = (C.c_int64 * n)()
        for i, k in enumerate(names):
            off, shp, ld = self.info[k]
            numel = int(np.prod(shp))
            r = shp[0] if (len(shp) == 2 and ld != numel) else 1
            offs[i], rows[i], cols[i], lds[i] = off, r, numel // r, ld if r > 1 else numel
        o = C.c_void_p()
        self.ctx.check(self.lib.ts_optim_create(self.ctx.h, n, offs, rows, cols, lds, self.n, C.byref(o)))
        return o

    def ensure_workspace(self, *key):
        if self.ws_key == key:
            return
        need = int(self._f("workspace_bytes")(self.h, *key))
        if need < 0:
            raise _lib.TethysError(-2, f"unsupported shape {key}: {self.lib.ts_last_error(self.ctx.h).decode()}")
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.ctx.check(self._f("bind")(self.h, ptr(self.params), ptr(self.grads), ptr(self.params_lp), ptr(self.workspace),
                                      self.workspace.numel()))
        self.ws_key = key

    def sync_weights(self):
        if not self.weights_synced and self.workspace is not None:
            self.ctx.check(self._f("sync_compute_weights")(self.h, stream_ptr()))
            self.weights_synced = True

    def buffer(self, name):
        p, dt, nd = C.c_void_p(), C.c_int32(), C.c_int32()
        shape = (C.c_int64 * 4)()
        self.ctx.check(self._f("get_buffer")(self.h, name.encode(), C.byref(p), C.byref(dt), C.byref(nd), shape))
        return view_from_ptr(p.value, [shape[i] for i in range(nd.value)], dt.value, self.device)

    def backward(self, stage_from=0, stage_to=10 ** 6):
        self.ctx.check(self._f("backward")(self.h, int(stage_from), int(stage_to), stream_ptr()))

    # -- bf16 gradient buckets for the all-reduce ("perf mode") -----------------------------------------------
    def ar_bf16(self):
        """bf16 buckets when the model computes in bf16 (its gradients carry bf16-level error already); fp32 parity mode
        always reduces in fp32. TETHYS_AR_DTYPE=fp32 forces fp32 buckets."""
        return self.precision == _lib.TS_BF16 and os.environ.get("TETHYS_AR_DTYPE", "bf16").lower() not in ("fp32", "float32", "f32")

    def grads_lp(self):
        if getattr(self, "_grads_lp", None) is None:
            st = getattr(self, "_comm_strategy", None)
            self._grads_lp = (st.alloc(self.n, torch.bfloat16) if st is not None
                              else torch.zeros(self.n, dtype=torch.bfloat16, device=self.device))
        return self._grads_lp

    def use_comm_buffers(self, strategy):
        """Move the gradient arenas (the buffers that cross NVLink) into memory registered with the native communicator, so that
        NCCL reduces them in place (NVLS / zero-copy) instead of staging through its own buffers. Idempotent."""
        if getattr(strategy, "comm", None) is None or getattr(self, "_comm_strategy", None) is strategy:
            return
        self._comm_strategy = strategy
        self._grads_lp = None
        if not self.ar_bf16():          # fp32 buckets: the fp32 gradient arena itself is reduced
            self.grads = strategy.alloc(self.n, torch.float32)
            self.ws_key = None          # re-bind the program to the new arena at the next ensure_workspace

    def pack_grads(self, a0=0, a1=None, scale=None):
        """grads_lp[a0:a1] = bf16(grads[a0:a1] * scale) — scale: device scalar or None."""
        a1 = self.n if a1 is None else a1
        g16 = self.grads_lp()
        self.ctx.check(self.lib.ts_grad_pack_bf16(self.ctx.h, ptr(self.grads[a0:a1]), ptr(g16[a0:a1]), a1 - a0, ptr(scale), stream_ptr()))
        return g16[a0:a1]

    def unpack_grads(self, a0=0, a1=None):
        a1 = self.n if a1 is None else a1
        self.ctx.check(self.lib.ts_grad_unpack_bf16(self.ctx.h, ptr(self.grads_lp()[a0:a1]), ptr(self.grads[a0:a1]), a1 - a0, stream_ptr()))

    def backward_allreduce_overlapped(self, strategy, bucket_elems=16 * 1024 * 1024):
        """K21: backward stage by stage; as soon as the arena prefix of a group of stages is final (the arena is laid out
        in backward-completion order, ts_*_stage_end) its bucket is all-reduced (SUM) asynchronously: NCCL's stream is
        ordered after the compute stream at the point of the call, so the reduction of bucket i runs while the later
        stages of backward are still computing. Returns after queuing a wait for every bucket on the compute stream."""
        ends = self.stage_ends
        works = []
        start = 0
        lp = strategy.dist is not None and self.ar_bf16()
        for s, end in enumerate(ends):
            self.backward(s, s)
            last = s == len(ends) - 1
            if strategy.dist is not None and (end - start >= bucket_elems or last) and end > start:
                bucket = self.pack_grads(start, end) if lp else self.grads[start:end]
                strategy.all_reduce_async_(bucket)
                works.append(1)
                start = end
        if works:
            strategy.join_async()
        if lp and works:
            self.unpack_grads()
        return len(works)


def rendezvous_from_tf_config(tf_config=None):
    """The replica group a TFJob pod belongs to, from the TF_CONFIG the operator injects (W:1037-1040, job_name.py:3-13;
    sample_tfjobs/*.yaml: one CHIEF + WORKER pods, one GPU and one `python speech_jobs/*_dist.py` process each):
        {"cluster": {"chief": ["host:2222"], "worker": ["host:2222", ...]}, "task": {"type": "worker", "index": 0}}
    -> (world, rank, master_host, master_port) with MultiWorkerMirroredStrategy's replica order (chief first, then the workers by
    index; `ps` / `evaluator` tasks do not train) and the chief's own TFJob port as the rendezvous address — nothing else is
    listening there, TensorFlow's gRPC server being gone. None when TF_CONFIG is absent or names a single task (W:1047 then
    degenerates to one worker)."""
    import json

    if tf_config is None:
        tf_config = os.environ.get("TF_CONFIG")
    if isinstance(tf_config, str):
        tf_config = json.loads(tf_config or "{}")
    if not tf_config:
        return None
    cluster = {str(k).lower(): list(v) for k, v in (tf_config.get("cluster") or {}).items()}
    task = tf_config.get("task") or {}
    order = [(job, i, addr) for job in ("chief", "master", "worker") for i, addr in enumerate(cluster.get(job, []))]
    if len(order) <= 1:
        return None
    me = (str(task.get("type", "")).lower(), int(task.get("index", 0)))
    ranks = [r for r, (job, i, _) in enumerate(order) if (job, i) == me]
    if not ranks:
        raise ValueError(f"TF_CONFIG task {me} is not a training task of the cluster {sorted(cluster)}")
    host, _, port = order[0][2].rpartition(":")
    return len(order), ranks[0], host, int(port)


class Strategy:
    """Stand-in for tf.distribute.MultiWorkerMirroredStrategy (W:1047, V:1473): one process per GPU. torch.distributed is the
    rendezvous (and the whole collective layer on CPU / gloo, for the host-logic tests); on GPUs the data path is the library's
    own communicator (ts_comm_*, csrc/comm.cu): NCCL called directly on the compute stream — capturable into the step's CUDA
    graph — over ncclMemAlloc-registered gradient arenas. TETHYS_NATIVE_COMM=0 keeps torch.distributed's NCCL for A/B runs.
    With a single process it degenerates to one replica."""

    def __init__(self, backend=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        init_method = None
        if "WORLD_SIZE" not in os.environ:
            # launched the reference's way — `python speech_jobs/*_dist.py` in every TFJob pod, no torchrun: the replica group
            # comes from TF_CONFIG like MultiWorkerMirroredStrategy's own cluster resolver (W:1037-1047)
            rdv = rendezvous_from_tf_config()
            if rdv is not None:
                self.world, self.rank, host, port = rdv
                init_method = f"tcp://{host}:{port}"
        self.dist = None
        self.comm = None            # ts_comm handle
        self.comm_stream = None     # side stream for collectives that overlap the backward pass
        self._ctx = None
        if self.world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                if backend is None:
                    backend = "nccl" if torch.cuda.is_available() else "gloo"
                if backend == "nccl":
                    torch.cuda.set_device(self.local_rank)
                if init_method is not None:
                    dist.init_process_group(backend=backend, init_method=init_method, rank=self.rank, world_size=self.world)
                else:
                    dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.dist = dist
            if torch.cuda.is_available() and dist.get_backend() == "nccl" and os.environ.get("TETHYS_NATIVE_COMM", "1") != "0":
                self._init_native()

    # -- native communicator -----------------------------------------------------------------------------------------------
    def _init_native(self):
        ctx = _lib.context(torch.cuda.current_device())
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            ctx.check(ctx.lib.ts_comm_unique_id(ctx.h, uid))
        box = [bytes(uid.raw)]
        self.dist.broadcast_object_list(box, src=0)          # the only thing torch.distributed carries: 128 bytes, once
        h = C.c_void_p()
        ctx.check(ctx.lib.ts_comm_init(ctx.h, box[0], self.world, self.rank, C.byref(h)))
        self.comm, self._ctx = h, ctx
        self.comm_stream = torch.cuda.Stream()

    def alloc(self, numel, dtype):
        """A flat device tensor in communicator-registered memory (ncclMemAlloc + ncclCommRegister); owned by the communicator."""
        p = C.c_void_p()
        esz = 4 if dtype == torch.float32 else 2
        self._ctx.check(self._ctx.lib.ts_comm_alloc(self.comm, int(numel) * esz, C.byref(p)))
        return view_from_ptr(p.value, (int(numel),), _lib.TS_F32 if dtype == torch.float32 else _lib.TS_BF16,
                             torch.device("cuda", torch.cuda.current_device()))

    def _native_all_reduce(self, t, premul=None):
        if os.environ.get("TETHYS_SKIP_AR") == "1":      # timing experiments only (tools/comm_bench.py): numerically wrong
            return
        dt = _lib.TS_F32 if t.dtype == torch.float32 else _lib.TS_BF16
        self._ctx.check(self._ctx.lib.ts_comm_allreduce_bucket(self.comm, ptr(t), t.numel(), dt, ptr(premul), stream_ptr()))

    def all_reduce_async_(self, t, premul=None, pre=None):
        """SUM all-reduce of `t` that may run underneath the kernels issued after it: enqueued on the communicator's side stream,
        ordered after everything already on the compute stream (a fork that a CUDA-graph capture records as such).
        join_async() makes the compute stream wait for all of them. `pre` (optional callable) runs on the same side stream right
        before the collective — the fp32 -> bf16 packing of the bucket, so that it too stays off the compute stream."""
        if self.comm is None:
            if pre is not None:
                pre()
            self._works = getattr(self, "_works", [])
            self._works.append(self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, async_op=True))
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            if pre is not None:
                pre()
            self._native_all_reduce(t, premul=premul)

    def async_done_event(self):
        """An event that fires when everything queued so far by all_reduce_async_ has finished (per-bucket joins)."""
        if self.comm is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self.comm_stream)
        return ev

    def join_async(self):
        if self.comm is None:
            for w in getattr(self, "_works", []):
                w.wait()
            self._works = []
            return
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def check(self):
        """ncclCommGetAsyncError of the native communicator (raises TethysError TS_ENCCL)."""
        if self.comm is not None:
            self._ctx.check(self._ctx.lib.ts_comm_check(self.comm))

    def info(self):
        if self.comm is None:
            return None
        a, b, v, r = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ctx.check(self._ctx.lib.ts_comm_info(self.comm, C.byref(a), C.byref(b), C.byref(v), C.byref(r)))
        return {"nranks": a.value, "rank": b.value, "nccl_version": v.value, "registered_buffers": r.value}

    def allreduce_description(self, prog):
        what = "bf16 gradient buckets (fp32 master weights / Adam state)" if prog.ar_bf16() else "fp32 gradient buckets"
        if self.comm is None:
            return what + "; torch.distributed NCCL, eager between graph segments"
        i = self.info()
        return (what + f"; ts_comm (NCCL {i['nccl_version']} called directly, captured in the step graph, "
                f"{i['registered_buffers']} ncclMemAlloc-registered arenas)")

    def shutdown(self):
        """Frees the registered arenas and destroys the communicator. Every CUDA graph that captured one of its collectives must
        have been destroyed first (NCCL requirement); a process that is about to exit can simply skip this."""
        if self.comm is not None:
            torch.cuda.synchronize()
            self._ctx.lib.ts_comm_finalize(self.comm)
            self.comm = None

    @property
    def num_replicas_in_sync(self):
        return self.world

    def scope(self):
        import contextlib

        return contextlib.nullcontext()

    def run(self, fn, args=()):
        return fn(*args)

    def reduce(self, op, value, axis=None):
        """strategy.reduce(SUM, per_replica_losses, axis=None) — W:848, V:1260."""
        if self.dist is None:
            return value
        if self.comm is not None and isinstance(value, torch.Tensor) and value.is_cuda:
            t = value.detach().float().reshape(-1).clone()
            self._native_all_reduce(t)
            return t.reshape(value.shape)
        t = value if isinstance(value, torch.Tensor) else torch.tensor(float(value))
        t = t.detach().clone().float()
        if torch.cuda.is_available() and self.dist.get_backend() == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def all_reduce_premul_sum_(self, flat, scale):
        """SUM over replicas of scale_r * flat_r with the per-replica device scalar `scale` applied inside the collective
        (ncclRedOpCreatePreMulSum); falls back to an explicit multiply if this torch/NCCL build lacks it."""
        if self.dist is None:
            flat.mul_(scale)
            return
        if self.comm is not None and flat.is_cuda and flat.dtype == scale.dtype:
            self._native_all_reduce(flat, premul=scale)
            return
        try:
            op = self.dist._make_nccl_premul_sum(scale)
        except Exception:  # noqa: BLE001 — gloo, or a build without pre-multiplied sums
            flat.mul_(scale)
            op = self.dist.ReduceOp.SUM
        self.dist.all_reduce(flat, op=op)

    def all_reduce_sum_(self, flat, bucket_elems=None):
        """K21: in-place SUM all-reduce of a flat gradient arena, issued as a few large buckets."""
        if self.dist is None:
            return
        if self.comm is not None and flat.is_cuda:
            self._native_all_reduce(flat)
            return
        n = flat.numel()
        if not bucket_elems or bucket_elems >= n:
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM)
            return
        works = []
        for s in range(0, n, bucket_elems):
            works.append(self.dist.all_reduce(flat[s:s + bucket_elems], op=self.dist.ReduceOp.SUM, async_op=True))
        for w in works:
            w.wait()

    def broadcast_(self, flat, src=0):
        """K23: weights created under strategy.scope() are mirrored from the chief (W:896-898, V:1266-1268)."""
        if self.comm is not None and flat.is_cuda:
            dt = _lib.TS_F32 if flat.dtype == torch.float32 else _lib.TS_BF16
            self._ctx.check(self._ctx.lib.ts_comm_broadcast(self.comm, ptr(flat), flat.numel(), dt, int(src), stream_ptr()))
        elif self.dist is not None:
            self.dist.broadcast(flat, src=src)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


class ReduceOp:
    SUM = "SUM"


_default_strategy = None


def get_strategy():
    global _default_strategy
    if _default_strategy is None:
        _default_strategy = Strategy()
    return _default_strategy


class GradientList(list):
    """What model.gradient() returns: views into the gradient arena, remembering the owning model so that
    optimizer.apply_gradients(zip(gradients, variables)) can run the fused multi-tensor update."""
    owner = None


class Adam:
    """tf.keras.optimizers.Adam of Keras 2.10 (legacy OptimizerV2 formula) — W:901, V:1271-1275.
    State (m, v) lives in flat fp32 arenas shaped like the model's parameter arena."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.clipnorm = clipnorm
        self.iterations = 0
        self._state = {}
        self.device_step = False   # True inside a GraphedTrainStep: the step count comes from the library's device state

    def _bind(self, model):
        key = id(model)
        if key not in self._state:
            prog = model._prog
            st = {"m": torch.zeros_like(prog.params), "v": torch.zeros_like(prog.params), "optim": prog.make_optim()}
            self._state[key] = st
        return self._state[key]

    def local_clip(self, model, global_clip_norm):
        """Phase 1 of the distributed apply (V:1243): clip_by_global_norm on this replica's gradients, in place."""
        prog = model._prog
        st = self._bind(model)
        prog.ctx.check(prog.lib.ts_optim_clip_global(st["optim"], ptr(prog.grads), float(global_clip_norm), None, stream_ptr()))

    def local_clip_scale(self, model, global_clip_norm, out):
        """clip_by_global_norm's factor for this replica's gradients (V:1243) written to the device scalar `out`, not applied:
        the distributed step folds it into the all-reduce (NCCL pre-multiplied sum)."""
        prog = model._prog
        st = self._bind(model)
        prog.ctx.check(prog.lib.ts_optim_global_clip_scale(st["optim"], ptr(prog.grads), float(global_clip_norm), ptr(out), stream_ptr()))

    def update(self, model, grads_lp=None):
        """Phase 3 (after the all-reduce): per-variable clipnorm + Adam, advancing `iterations`. grads_lp: the all-reduced bf16
        gradient bucket (whole arena) — read as is (ts_optim_step_lp) instead of being unpacked into the fp32 arena first."""
        prog = model._prog
        st = self._bind(model)
        self.iterations += 1
        fn = prog.lib.ts_optim_step if grads_lp is None else prog.lib.ts_optim_step_lp
        prog.ctx.check(fn(st["optim"], ptr(prog.params), ptr(prog.grads if grads_lp is None else grads_lp), ptr(st["m"]), ptr(st["v"]),
                                              ptr(prog.params_lp), self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                                              0 if self.device_step else self.iterations, 0.0, float(self.clipnorm or 0.0), 0,
                                              stream_ptr()))
        prog.weights_synced = True

    def update_range(self, model, optim_handle):
        """clipnorm + Adam on the variables of one all-reduce bucket (its own ts_optim); does not touch `iterations` — the
        caller advances it once per step."""
        prog = model._prog
        st = self._bind(model)
        prog.ctx.check(prog.lib.ts_optim_step(optim_handle, ptr(prog.params), ptr(prog.grads), ptr(st["m"]), ptr(st["v"]),
                                              ptr(prog.params_lp), self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                                              0 if self.device_step else self.iterations + 1, 0.0, float(self.clipnorm or 0.0), 0,
                                              stream_ptr()))
        prog.weights_synced = True

    def apply_gradients(self, grads_and_vars, strategy=None, global_clip_norm=None, model=None, already_reduced=False):
        """All-reduce (SUM, un-normalised: App. A-13) the gradient arena across replicas, apply the per-variable
        clipnorm, then the Adam update. `global_clip_norm` fuses tf.clip_by_global_norm into the same pass when no
        all-reduce sits in between (single replica)."""
        if model is None:
            gl = grads_and_vars
            if not isinstance(gl, GradientList):
                pairs = list(grads_and_vars)
                gl = None
                for g, _ in pairs:
                    gl = getattr(g, "_ts_owner", None)
                    if gl is not None:
                        break
                model = gl
            else:
                model = gl.owner
        if model is None:
            raise ValueError("apply_gradients: gradients must come from model.gradient()")
        prog = model._prog
        st = self._bind(model)
        strategy = strategy or get_strategy()
        lib, ctx = prog.lib, prog.ctx
        fuse = 0
        gclip = 0.0
        if global_clip_norm:
            if strategy.num_replicas_in_sync > 1:
                ctx.check(lib.ts_optim_clip_global(st["optim"], ptr(prog.grads), float(global_clip_norm), None, stream_ptr()))
            else:
                fuse, gclip = 1, float(global_clip_norm)
        if strategy.num_replicas_in_sync > 1 and not already_reduced:
            if strategy.dist is not None and prog.ar_bf16():
                strategy.all_reduce_sum_(prog.pack_grads())
                prog.unpack_grads()
            else:
                strategy.all_reduce_sum_(prog.grads, bucket_elems=int(os.environ.get("TETHYS_AR_BUCKET_ELEMS", 0)))
        self.iterations += 1
        ctx.check(lib.ts_optim_step(st["optim"], ptr(prog.params), ptr(prog.grads), ptr(st["m"]), ptr(st["v"]),
                                    ptr(prog.params_lp), self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                                    0 if self.device_step else self.iterations, gclip, float(self.clipnorm or 0.0), fuse, stream_ptr()))
        prog.weights_synced = True  # the update refreshed the bf16 compute copy in the same pass


def native_step_args(model, optimizer, strategy=None, global_clip=0.0, dropout=True, seed=0):
    """ts_step_args for the composite C-ABI entries (ts_w2v_step / ts_whisper_step): the optimizer state this
    Adam object keeps for `model`, the replica group of `strategy` (its native communicator; None / one replica: no collective)
    and a device scalar for the step's return value. Returns (args, loss_out); keep both alive until the step has run."""
    prog = model._prog
    st = optimizer._bind(model)
    a = _lib.StepArgs()
    a.optim, a.adam_m, a.adam_v = st["optim"], ptr(st["m"]), ptr(st["v"])
    a.lr, a.beta1, a.beta2, a.eps = optimizer.learning_rate, optimizer.beta_1, optimizer.beta_2, optimizer.epsilon
    a.step = 0 if optimizer.device_step else optimizer.iterations + 1
    a.global_clip, a.clipnorm = float(global_clip or 0.0), float(optimizer.clipnorm or 0.0)
    a.seed, a.dropout = int(seed), 1 if dropout else 0
    comm = getattr(strategy, "comm", None) if strategy is not None else None
    if comm is not None:
        a.comm = comm
        if prog.ar_bf16():
            a.grads_bf16 = ptr(prog.grads_lp())
        if "scratch" not in st:
            st["scratch"] = torch.zeros(2, dtype=torch.float32, device=prog.device)
        a.scratch_dev = ptr(st["scratch"])
    if "loss_out" not in st:
        st["loss_out"] = torch.zeros(1, dtype=torch.float32, device=prog.device)
    a.loss_out_dev = ptr(st["loss_out"])
    return a, st["loss_out"]


class GraphedTrainStep:
    """One whole train step (forward, loss, backward, clip, Adam) captured in a CUDA graph and replayed: removes the ~430
    kernel-launch gaps of a step. The library's device-resident step state (ts_step_state_*) gives every replay fresh
    dropout masks and the right Adam bias correction; inputs are copied into static device buffers before each replay.

        graphed = GraphedTrainStep(lambda batch, aux: train_step(model, batch, optimizer, **aux), model, optimizer,
                                   example_batch, example_aux)
        loss = graphed(batch, aux)          # loss is a device scalar (a view of the workspace)

    With more than one replica the step functions capture the native communicator's collectives into the same graph
    (make_graphed_distributed_step in wav2vec2.py / whisper.py)."""

    def __init__(self, step_fn, model, optimizer, example_batch, example_aux=None, warmup=3):
        self.step_fn, self.model, self.opt = step_fn, model, optimizer
        prog = model._prog
        self.ctx = prog.ctx
        dev = prog.device
        self.static_batch = tuple(None if t is None else to_device(t, t.dtype if isinstance(t, torch.Tensor) else None, dev).clone()
                                  for t in example_batch)
        self.static_aux = {k: v.to(dev).clone() for k, v in (example_aux or {}).items()}
        # `warmup` eager passes allocate the workspace / optimizer state and set kernel attributes. They are REAL steps on the
        # example batch, so parameters, Adam moments, the iteration count and the dropout seed are snapshotted first and put
        # back afterwards: building the graph has no training side effect.
        st = optimizer._bind(model)
        snap = (prog.params.clone(), st["m"].clone(), st["v"].clone(), optimizer.iterations, getattr(model, "_step_seed", None))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(self.static_batch, self.static_aux)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if warmup:
            prog.params.copy_(snap[0]); st["m"].copy_(snap[1]); st["v"].copy_(snap[2])
            optimizer.iterations = snap[3]
            if snap[4] is not None:
                model._step_seed = snap[4]
            prog.weights_synced = False
            prog.sync_weights()
        self.ctx.check(self.ctx.lib.ts_step_state_set(self.ctx.h, 0, int(optimizer.iterations), stream_ptr()))
        optimizer.device_step = True
        l0 = self.ctx.lib.ts_launch_count(self.ctx.h)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.ctx.check(self.ctx.lib.ts_step_state_advance(self.ctx.h, stream_ptr()))
            self.loss = step_fn(self.static_batch, self.static_aux)
        optimizer.device_step = False
        optimizer.iterations -= 1              # the capture recorded a step but did not run it
        self.launches_per_step = int(self.ctx.lib.ts_launch_count(self.ctx.h) - l0)

    def __call__(self, batch, aux=None):
        for dst, src in zip(self.static_batch, batch):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        for k, v in (aux or {}).items():
            self.static_aux[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.opt.iterations += 1
        return self.loss


class GraphedSegments:
    """A train step as an alternating list of captured CUDA graphs and eager hooks (the NCCL collectives):
        plan = [("graph", fn), ("eager", fn), ("side_graph", fn), ...]
    Every fn takes no arguments and works on static device buffers; "graph" items are captured once (after `warmup` eager
    passes over the whole plan) and replayed on the current stream, "side_graph" items are captured and replayed on
    `side_stream` (work that may run underneath the main stream, e.g. the Adam update of an already reduced bucket),
    "eager" items are called as they are. This keeps the collectives outside the graphs — so a bucket's all-reduce still
    overlaps the next graph segment — while removing the launch gaps of the ~430 kernels of a step. The first graph
    segment must start with ts_step_state_advance (see GraphedTrainStep). `bump_iterations`: the plan's update items do
    not advance optimizer.iterations themselves (per-bucket updates), so the step does it once."""

    def __init__(self, plan, model, optimizer, warmup=3, side_stream=None, bump_iterations=False):
        self.plan, self.opt, self.bump = plan, optimizer, bump_iterations
        self.ctx = model._prog.ctx
        dev = model._prog.device
        self.side = side_stream

        def run_side(fn):
            with torch.cuda.stream(self.side):
                fn()

        for _ in range(warmup):
            for kind, fn in plan:
                if kind == "side_graph":
                    run_side(fn)
                else:
                    fn()
            if bump_iterations:
                optimizer.iterations += 1
        torch.cuda.synchronize(dev)
        self.ctx.check(self.ctx.lib.ts_step_state_set(self.ctx.h, 0, int(optimizer.iterations), stream_ptr()))
        torch.cuda.synchronize(dev)
        it0 = optimizer.iterations
        optimizer.device_step = True
        l0 = self.ctx.lib.ts_launch_count(self.ctx.h)
        self.items = []
        for kind, fn in plan:
            if kind in ("graph", "side_graph"):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                if kind == "graph":
                    self.items.append(g.replay)
                else:
                    self.items.append(lambda g=g: run_side(g.replay))
            else:
                self.items.append(fn)
        optimizer.device_step = False
        optimizer.iterations = it0              # captures record work, they do not run it
        self.launches_per_step = int(self.ctx.lib.ts_launch_count(self.ctx.h) - l0)
        torch.cuda.synchronize(dev)

    def __call__(self):
        for run in self.items:
            run()
        self.opt.iterations += 1


def clip_by_global_norm(gradients, clip_norm):
    """tf.clip_by_global_norm on the gradient arena, in place — V:1243. Returns (gradients, global_norm)."""
    model = gradients.owner
    prog = model._prog
    opt = prog.make_optim()
    norm = torch.zeros(1, device=prog.device)
    prog.ctx.check(prog.lib.ts_optim_clip_global(opt, ptr(prog.grads), float(clip_norm), ptr(norm), stream_ptr()))
    return gradients, norm
