"""Checkpoint save and restore from the flat arenas.

The reference writes `tf.train.Checkpoint(model=model, optimizer=optimizer).save(...)` every 50 steps and at the end of an
epoch (V:1286-1288, V:1341, V:1362; W:917-919, W:956) and `model.save_weights` once (W:1025); it never restores. Here the
whole state of a run — parameters, Adam moments, `optimizer.iterations`, the dropout step seed — is three contiguous fp32
arenas plus two integers, so a save is one device→pinned-host copy per arena followed by a sequential write, and a
restore is the mirror image. Variables are stored one by one under the reference's Keras variable paths (not as an
arena dump): the arena order is a backward-stage order private to the library and may change between versions.

File layout (little endian):
    8 B   magic  b"TSCKPT01"
    8 B   u64    header length H
    H B   JSON   {"meta": {...}, "tensors": {key: {"offset": o, "shape": [...], "dtype": "<f4"}}}
    ...   raw arrays, each starting at a 64-byte aligned `offset` counted from the end of the header
Keys: "model/<variable path>", "optimizer/m/<variable path>", "optimizer/v/<variable path>".
"""
import glob
import json
import os
import re
import struct

import numpy as np

MAGIC = b"TSCKPT01"
_ALIGN = 64


# ----------------------------------------------------------------------------------------------------------------------
# arena <-> named variables (pure numpy; `info` is ProgramBase.info: name -> (offset, shape, row stride))
# ----------------------------------------------------------------------------------------------------------------------
def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def gather_variables(arena, info, names=None):
    """name -> contiguous fp32 array cut out of a host copy of an arena. Column slices of fused blocks (q/k/v side by
    side, row stride ld != cols) are gathered row by row."""
    out = {}
    for name in (names if names is not None else info):
        off, shp, ld = info[name]
        n = _numel(shp)
        if len(shp) == 2 and ld != n:
            rows, cols = shp
            v = np.lib.stride_tricks.as_strided(arena[off:], shape=(rows, cols), strides=(ld * arena.itemsize, arena.itemsize))
            out[name] = np.ascontiguousarray(v)
        else:
            out[name] = np.ascontiguousarray(arena[off:off + n]).reshape(shp)
    return out


def scatter_variables(arena, info, tensors, strict=True):
    """Inverse of gather_variables: write named arrays into a host arena in place. Returns the list of names written.
    strict: every variable of `info` must be present with the right shape; otherwise missing ones are left untouched."""
    done = []
    for name, (off, shp, ld) in info.items():
        if name not in tensors:
            if strict:
                raise KeyError(f"checkpoint has no variable '{name}'")
            continue
        a = np.asarray(tensors[name], dtype=arena.dtype)
        if tuple(a.shape) != tuple(shp):
            raise ValueError(f"checkpoint variable '{name}' has shape {tuple(a.shape)}, the model expects {tuple(shp)}")
        n = _numel(shp)
        if len(shp) == 2 and ld != n:
            rows, cols = shp
            v = np.lib.stride_tricks.as_strided(arena[off:], shape=(rows, cols), strides=(ld * arena.itemsize, arena.itemsize))
            v[...] = a
        else:
            arena[off:off + n] = a.reshape(-1)
        done.append(name)
    if strict:
        extra = [k for k in tensors if k not in info]
        if extra:
            raise KeyError(f"checkpoint has variables the model does not: {extra[:4]}{' ...' if len(extra) > 4 else ''}")
    return done


# ----------------------------------------------------------------------------------------------------------------------
# file format
# ----------------------------------------------------------------------------------------------------------------------
def write_file(path, tensors, meta):
    """tensors: key -> numpy array. Written to `path + '.tmp'` and renamed, so a crash never leaves a torn checkpoint."""
    index, pos = {}, 0
    for k, a in tensors.items():
        pos = (pos + _ALIGN - 1) // _ALIGN * _ALIGN
        index[k] = {"offset": pos, "shape": [int(s) for s in a.shape], "dtype": a.dtype.str}
        pos += a.nbytes
    header = json.dumps({"meta": meta, "tensors": index}).encode()
    base = len(MAGIC) + 8 + len(header)
    pad0 = (-base) % _ALIGN
    header += b" " * pad0
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        pos = 0
        for k, a in tensors.items():
            o = index[k]["offset"]
            if o > pos:
                f.write(b"\0" * (o - pos))
            f.write(memoryview(np.ascontiguousarray(a).reshape(-1)).cast("B"))
            pos = o + a.nbytes
    os.replace(tmp, path)
    return path


def read_file(path, keys=None):
    """-> (meta, {key: array}); arrays are read-only views of one memory map. keys: optional filter (callable or set)."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a tethys checkpoint")
        (hlen,) = struct.unpack("<Q", f.read(8))
        hdr = json.loads(f.read(hlen).decode())
    base = len(MAGIC) + 8 + hlen
    size = os.path.getsize(path)
    mm = np.memmap(path, dtype=np.uint8, mode="r") if size > base else np.zeros(0, np.uint8)
    out = {}
    for k, d in hdr["tensors"].items():
        if keys is not None and not (keys(k) if callable(keys) else k in keys):
            continue
        dt = np.dtype(d["dtype"])
        n = _numel(d["shape"]) * dt.itemsize
        o = base + d["offset"]
        if o + n > size:
            raise ValueError(f"{path}: truncated (tensor '{k}' ends at byte {o + n}, file has {size})")
        out[k] = mm[o:o + n].view(dt).reshape(d["shape"])
    return hdr["meta"], out


# ----------------------------------------------------------------------------------------------------------------------
# TF object-graph keys (export side): tf.train.Checkpoint(model=…) names a Keras variable by the attribute path from the
# root object. The reference's attribute names are the ones in its constructors (V:746-766, V:464-546, W:470-545 …).
# TensorFlow is not installable here, so this mapping is written from the reference source and is NOT verified against a
# TF-written checkpoint; it is used only for `export_tf_names`.
# ----------------------------------------------------------------------------------------------------------------------
def tf_object_key(variable_path, root="model"):
    return root + "/" + variable_path.replace(".", "/") + "/.ATTRIBUTES/VARIABLE_VALUE"


def export_tf_names(model):
    """variable path -> TF object-graph key for every trainable variable (same order as model.trainable_variables)."""
    return {n: tf_object_key(n) for n in model.variable_names}


# ----------------------------------------------------------------------------------------------------------------------
# model / optimizer level
# ----------------------------------------------------------------------------------------------------------------------
def _to_host(t):
    """One device→pinned-host copy of a whole arena (a 369 MB Wav2Vec2-base arena moves in ~8 ms over PCIe 5)."""
    import torch
    if not t.is_cuda:
        return t.detach().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _from_host(a, t):
    import torch
    h = torch.from_numpy(a)
    if t.is_cuda:
        h = h.pin_memory()
    t.copy_(h, non_blocking=True)
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


def save(path, model, optimizer=None, extra_meta=None):
    """Write the model's variables (and, with `optimizer`, the Adam moments + iteration count) to `path`."""
    prog = model._prog
    names = list(model.variable_names)
    tensors = {"model/" + k: v for k, v in gather_variables(_to_host(prog.params), prog.info, names).items()}
    meta = {"format": 1, "kind": type(model).__name__, "precision": "bf16" if prog.params_lp is not None else "fp32",
            "step_seed": int(getattr(model, "_step_seed", 0)), "variables": names}
    import ctypes as C
    salt, dstep = C.c_uint64(), C.c_int64()
    prog.ctx.check(prog.lib.ts_step_state_get(prog.ctx.h, C.byref(salt), C.byref(dstep)))
    meta["device_salt"] = int(salt.value)          # the dropout salt a CUDA-graph replay of the step reads (ts_step_state_*)
    if optimizer is not None:
        st = optimizer._bind(model)
        for slot in ("m", "v"):
            for k, v in gather_variables(_to_host(st[slot]), prog.info, names).items():
                tensors[f"optimizer/{slot}/{k}"] = v
        meta["optimizer"] = {"iterations": int(optimizer.iterations), "learning_rate": optimizer.learning_rate,
                             "beta_1": optimizer.beta_1, "beta_2": optimizer.beta_2, "epsilon": optimizer.epsilon,
                             "clipnorm": optimizer.clipnorm}
    if extra_meta:
        meta.update(extra_meta)
    return write_file(path, tensors, meta)


def restore(path, model, optimizer=None, strict=True):
    """Load `path` into the model's arenas (and the optimizer's, when both the file and the call have one). The bf16
    compute copy is refreshed on the next forward. Returns the checkpoint's meta dict."""
    prog = model._prog
    meta, tensors = read_file(path)
    host = _to_host(prog.params).copy()
    scatter_variables(host, prog.info, {k[6:]: v for k, v in tensors.items() if k.startswith("model/")}, strict=strict)
    _from_host(host, prog.params)
    prog.weights_synced = False
    if "step_seed" in meta and hasattr(model, "_step_seed"):
        model._step_seed = int(meta["step_seed"])
    if optimizer is not None:
        if "optimizer" not in meta:
            if strict:
                raise KeyError(f"{path} holds no optimizer state")
        else:
            st = optimizer._bind(model)
            for slot in ("m", "v"):
                pre = f"optimizer/{slot}/"
                host = _to_host(st[slot]).copy()
                scatter_variables(host, prog.info, {k[len(pre):]: v for k, v in tensors.items() if k.startswith(pre)}, strict=strict)
                _from_host(host, st[slot])
            optimizer.iterations = int(meta["optimizer"]["iterations"])
    # A step captured in a CUDA graph (GraphedTrainStep / GraphedSegments) reads the Adam step count and the dropout salt from
    # the library's device state, not from the host objects: refresh it unconditionally, so that a restore AFTER the graphs were
    # built continues with the right bias correction and the masks the straight run would have drawn.
    from .runtime import stream_ptr
    step_now = int(optimizer.iterations) if optimizer is not None else 0
    prog.ctx.check(prog.lib.ts_step_state_set(prog.ctx.h, int(meta.get("device_salt", step_now)), step_now, stream_ptr()))
    return meta


class Checkpoint:
    """tf.train.Checkpoint(model=…, optimizer=…) stand-in (V:1286-1288): `save(file_prefix)` numbers the files like TF's
    save counter (`<prefix>-<n>`), `restore(path)` is the restore the reference never calls."""

    def __init__(self, model=None, optimizer=None):
        self.model, self.optimizer = model, optimizer
        self.save_counter = 0

    def save(self, file_prefix):
        self.save_counter += 1
        d = os.path.dirname(file_prefix)
        if d:
            os.makedirs(d, exist_ok=True)
        path = f"{file_prefix}-{self.save_counter}.tsckpt"
        save(path, self.model, self.optimizer, {"save_counter": self.save_counter})
        return path

    def write(self, path):
        return save(path, self.model, self.optimizer)

    def restore(self, path, strict=True):
        meta = restore(path, self.model, self.optimizer, strict=strict)
        self.save_counter = int(meta.get("save_counter", self.save_counter))
        return meta

    read = restore


def latest_checkpoint(checkpoint_dir, prefix=None):
    """tf.train.latest_checkpoint: the file with the highest save counter (`<prefix>-<n>.tsckpt`, what Checkpoint.save writes)
    in `checkpoint_dir`; files without a counter (written by plain `save(path, ...)`) rank by modification time below any
    numbered file. None if the directory holds no checkpoint."""
    best, best_key = None, None
    for p in glob.glob(os.path.join(checkpoint_dir, "*.tsckpt")):
        if prefix and not os.path.basename(p).startswith(prefix):
            continue
        m = re.search(r"-(\d+)\.tsckpt$", p)
        key = (1, int(m.group(1)), os.path.getmtime(p)) if m else (0, 0, os.path.getmtime(p))
        if best_key is None or key > best_key:
            best, best_key = p, key
    return best
