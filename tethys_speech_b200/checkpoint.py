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
# TF object-graph keys (export side). tf.train.Checkpoint(model=model, optimizer=optimizer) (V:1303, W:919) names a variable by
# the FIRST breadth-first path of Python attribute names from the root: attribute names of sub-layers, list indices
# (`layers/0`), `layer_with_weights-<k>` for the layers of a tf.keras.Sequential (the conv blocks of V:239-268), the
# add_weight / attribute name of the variable itself, then "/.ATTRIBUTES/VARIABLE_VALUE". The library's variable paths differ
# from the reference's attribute structure in a few places (`fe.` = wav2vec2.feature_extractor, the Sequential conv blocks, the
# task heads that live on the outer model), so the mapping is a short rule table. It is pinned by walking the reference's OWN
# model objects (its unmodified constructors, running on the test suite's TensorFlow stand-in) with those naming rules:
# tests/test_checkpoint.py::test_tf_object_keys_follow_the_reference_object_graph. The rules themselves are restated from
# TensorFlow 2.10's trackable code (TensorFlow is not installable here): the keys are not checked against a TF-written file.
# ----------------------------------------------------------------------------------------------------------------------
_TF_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def _tf_object_path(variable_path, kind):
    if kind == "whisper":
        # WhisperForConditionalGeneration: self.model = WhisperModel(config) (W:541), self.lm_head (W:545)
        if variable_path.startswith("lm_head."):
            return variable_path.replace(".", "/")
        return "model/" + variable_path.replace(".", "/")
    # Wav2Vec2ForPreTraining / ForCTC / ForSequenceClassification: self.wav2vec2 = Wav2Vec2Model(config) (V:832, V:944, V:1008)
    m = re.match(r"fe\.conv(\d+)\.(gn\.)?(\w+)$", variable_path)
    if m:       # conv_layers[i] = Sequential([Conv1D, GroupNormalization, Activation]) (V:239-268)
        return f"wav2vec2/feature_extractor/conv_layers/{m.group(1)}/layer_with_weights-{1 if m.group(2) else 0}/{m.group(3)}"
    if variable_path.startswith("fe.pos_conv."):
        return "wav2vec2/feature_extractor/pos_conv_embed/" + variable_path[len("fe.pos_conv."):]
    if variable_path.startswith("fe."):
        return "wav2vec2/feature_extractor/" + variable_path[3:].replace(".", "/")
    if variable_path.startswith("lm_head."):                 # Wav2Vec2ForCTC.lm_head (V:948)
        return variable_path.replace(".", "/")
    if variable_path.startswith("classifier_proj."):         # Wav2Vec2ForSequenceClassification.projector (V:1012)
        return "projector/" + variable_path[len("classifier_proj."):]
    if variable_path.startswith("classifier."):              # .classifier (V:1015)
        return variable_path.replace(".", "/")
    return "wav2vec2/" + variable_path.replace(".", "/")


def tf_object_key(variable_path, kind="wav2vec2", root="model"):
    """Checkpoint key of one model variable in a file written by the reference's tf.train.Checkpoint(model=...)."""
    return root + "/" + _tf_object_path(variable_path, kind) + _TF_SUFFIX


def tf_slot_key(variable_path, slot, kind="wav2vec2", root="model", optimizer="optimizer"):
    """Key of an optimizer slot variable (Adam's "m" / "v") of the same checkpoint: TF hangs slots under the variable they
    belong to, `<variable>/.OPTIMIZER_SLOT/<optimizer attribute>/<slot>`."""
    return f"{root}/{_tf_object_path(variable_path, kind)}/.OPTIMIZER_SLOT/{optimizer}/{slot}{_TF_SUFFIX}"


def _kind_of(model):
    return "whisper" if type(model).__name__.startswith("Whisper") else "wav2vec2"


def export_tf_names(model, with_optimizer=False):
    """variable path -> TF object-graph key for every trainable variable (same order as model.trainable_variables); with
    `with_optimizer` also "optimizer/m/<path>", "optimizer/v/<path>" (this file format's keys) -> the TF slot keys and
    "optimizer/iterations" -> Keras' `optimizer/iter`."""
    kind = _kind_of(model)
    out = {n: tf_object_key(n, kind) for n in model.variable_names}
    if with_optimizer:
        for n in model.variable_names:
            out["optimizer/m/" + n] = tf_slot_key(n, "m", kind)
            out["optimizer/v/" + n] = tf_slot_key(n, "v", kind)
        out["optimizer/iterations"] = "optimizer/iter" + _TF_SUFFIX
    return out


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
    # every stored tensor's name in a checkpoint the reference's tf.train.Checkpoint(model=, optimizer=) writes (see tf_object_key)
    tfk = export_tf_names(model, with_optimizer=optimizer is not None)
    meta["tf_keys"] = {("model/" + k if not k.startswith("optimizer/") else k): v for k, v in tfk.items() if k != "optimizer/iterations"}
    if extra_meta:
        meta.update(extra_meta)
    return write_file(path, tensors, meta)


def tf_named_tensors(path):
    """The tensors of a checkpoint file keyed by their TF object-graph names ({tf key: array}; plus `optimizer/iter` when the
    file holds optimizer state) — what a TF-side importer assigns to a freshly built reference model
    (`tf.train.load_variable`-style names; SURVEY f-3 "TF-checkpoint-name-compatible export")."""
    meta, tensors = read_file(path)
    tfk = meta.get("tf_keys")
    if tfk is None:
        raise KeyError(f"{path} was written before TF names were recorded; re-save it")
    out = {tfk[k]: v for k, v in tensors.items()}
    if "optimizer" in meta:
        out["optimizer/iter" + _TF_SUFFIX] = np.asarray(int(meta["optimizer"]["iterations"]), dtype=np.int64)
    return out


def assign_to_keras(model, path_or_named, optimizer=None, root="model"):
    """The import side of the TF-name export, for a process that runs the REFERENCE's Keras model: resolve every TF
    object-graph key of a checkpoint file (or of a {tf key: array} dict from tf_named_tensors) on `model` by walking the same
    attribute path TensorFlow's tracking walks — attribute names, list indices, `layer_with_weights-<k>` of a Sequential — and
    `assign` the stored value to the tf.Variable found there. With a Keras optimizer (`get_slot`, `iterations`) the Adam
    slots and the iteration count are assigned too. Pure Python: no TensorFlow import here; anything that exposes the
    reference's attribute structure works (the tests run it on the reference's own classes). Returns the number of
    model variables assigned."""
    named = tf_named_tensors(path_or_named) if isinstance(path_or_named, str) else dict(path_or_named)

    def resolve(obj, segs):
        for seg in segs:
            m = re.match(r"layer_with_weights-(\d+)$", seg)
            if m:
                layers = [l for l in obj.layers if (getattr(l, "weights", None) or getattr(l, "trainable_variables", None))]
                obj = layers[int(m.group(1))]
            elif seg.isdigit() and not hasattr(obj, seg):
                obj = obj[int(seg)]
            else:
                obj = getattr(obj, seg)
        return obj

    done = 0
    for key, value in named.items():
        if not key.endswith(_TF_SUFFIX):
            continue
        body = key[:-len(_TF_SUFFIX)]
        if body == "optimizer/iter":
            if optimizer is not None and hasattr(getattr(optimizer, "iterations", None), "assign"):
                optimizer.iterations.assign(int(value))
            continue
        slot = None
        if "/.OPTIMIZER_SLOT/" in body:
            body, tail = body.split("/.OPTIMIZER_SLOT/")
            slot = tail.split("/")[-1]
        segs = body.split("/")
        if segs[0] != root:
            continue
        var = resolve(model, segs[1:])
        if slot is None:
            var.assign(np.array(value))
            done += 1
        elif optimizer is not None and hasattr(optimizer, "get_slot"):
            optimizer.get_slot(var, slot).assign(np.array(value))
    return done


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
