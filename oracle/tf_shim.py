"""TEST INFRASTRUCTURE ONLY — a minimal `tensorflow` stand-in on PyTorch-CPU, just large enough to IMPORT AND RUN THE
UNMODIFIED reference scripts (/root/reference/speech_jobs/*.py) in this container, where TensorFlow is not installable.

Why: `oracle/*_oracle.py` restate the reference's models by hand; what can go wrong in a restatement is STRUCTURE (layer
order, which tensor feeds the quantiser, dropout sites, residual wiring, label shifts, the step's clip/reduce order). Running the
reference's own classes and step functions on this shim and diffing them against the oracle (oracle/ref_runner.py,
tests/test_reference_pinning.py, tests/golden/make_ref_golden.py) removes that risk: after it, only the OP SEMANTICS below —
each a few lines, written independently of oracle/tf_ops.py and citing the TF/Keras 2.10 behaviour they follow
— remain restated. Nothing here is imported by the product.

Design: tf tensors ARE torch tensors (autograd gives GradientTape); `tf.shape` returns Python ints; `tf.float32` resolves to
FLOATX (float32 by default, float64 for the 1e-10 pinning runs); Keras' layer machinery (lazy build, attribute tracking order
of trainable_variables, propagation of `training` through the call context) is reproduced because the reference relies on it
(e.g. project_hid/project_q are called without `training=` inside a training=True call, V:854-857).
"""
import contextlib
import inspect
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

FLOATX = torch.float32          # what `tf.float32` means; tests switch it to float64 (set_floatx)
_RNG = torch.Generator().manual_seed(0)
RANDOM_LOG = []                 # every tf.random.* draw, in order: (kind, tensor) — lets tests inject the same draws into the oracle


def set_floatx(dt):
    global FLOATX
    FLOATX = dt


def seed(s):
    _RNG.manual_seed(int(s))
    RANDOM_LOG.clear()


class _DT:
    def __init__(self, name, torch_dtype=None, size=4):
        self.name, self._t, self.size = name, torch_dtype, size

    def __repr__(self):
        return f"tf.{self.name}"


float32 = _DT("float32", None, 4)
float64 = _DT("float64", torch.float64, 8)
float16 = _DT("float16", torch.float16, 2)
int32 = _DT("int32", torch.int32, 4)
int64 = _DT("int64", torch.int64, 8)
bool_ = _DT("bool", torch.bool, 1)


def _dt(d, default=None):
    if d is None:
        return default
    if isinstance(d, _DT):
        return FLOATX if d is float32 else d._t
    if isinstance(d, torch.dtype):
        return d
    if d in ("float32", float):
        return FLOATX
    if d in ("int32", int):
        return torch.int32
    if d == "int64":
        return torch.int64
    raise TypeError(f"tf_shim: unknown dtype {d!r}")


def _is_t(x):
    return isinstance(x, torch.Tensor)


def _t(x, dtype=None):
    """convert_to_tensor: python floats -> FLOATX, python ints -> int32, numpy float64 stays unless dtype is given."""
    if _is_t(x):
        return x if dtype is None else x.to(_dt(dtype))
    if isinstance(x, (list, tuple)) and any(_is_t(e) for e in x):
        return torch.stack([_t(e, dtype) for e in x])
    a = np.asarray(x)
    if dtype is not None:
        if dtype is float32 and a.dtype.kind == "f":
            # a float32 CONSTANT holds float32-rounded values in TF (e.g. the sinusoid table, W:66) even when this shim
            # evaluates the graph in float64
            return torch.as_tensor(a.astype(np.float32)).to(FLOATX)
        return torch.as_tensor(a).to(_dt(dtype))
    if a.dtype.kind == "f":
        return torch.as_tensor(a).to(FLOATX if a.dtype != np.float64 or isinstance(x, (float, list, tuple)) else torch.float64)
    if a.dtype.kind in "iu":
        return torch.as_tensor(a).to(torch.int32 if isinstance(x, (int, list, tuple)) else torch.as_tensor(a).dtype)
    return torch.as_tensor(a)


# ---------------------------------------------------------------------------------------------------------------------
# basic ops
# ---------------------------------------------------------------------------------------------------------------------
def shape(x):
    return tuple(int(s) for s in _t(x).shape)


def constant(v, dtype=None, shape=None):
    t = _t(v, dtype)
    return t.reshape(shape) if shape is not None else t


def convert_to_tensor(v, dtype=None):
    return _t(v, dtype)


def cast(x, dtype):
    if not _is_t(x) and isinstance(x, (int, float, bool, np.integer, np.floating)):
        return torch.tensor(x, dtype=_dt(dtype))
    return _t(x).to(_dt(dtype))


def zeros(shp, dtype=float32):
    return torch.zeros(tuple(int(s) for s in shp), dtype=_dt(dtype))


def ones(shp, dtype=float32):
    if isinstance(shp, (int, np.integer)):
        shp = (shp,)
    return torch.ones(tuple(int(s) for s in shp), dtype=_dt(dtype))


def fill(dims, value):
    return torch.full(tuple(int(s) for s in dims), value, dtype=(FLOATX if isinstance(value, float) else (torch.bool if isinstance(value, bool) else torch.int32)))


def zeros_like(x, dtype=None):
    return torch.zeros_like(x, dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return torch.ones_like(x, dtype=_dt(dtype))


def range_(*args, dtype=None):
    if any(_is_t(a) for a in args):
        args = [int(a) for a in args]
    isf = any(isinstance(a, float) for a in args)
    return torch.arange(*args, dtype=_dt(dtype, FLOATX if isf else torch.int32))


def reshape(x, shp):
    return x.reshape(tuple(int(s) for s in shp))


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim())))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def pad(x, paddings, mode="CONSTANT", constant_values=0):
    flat = []
    for lo, hi in reversed([tuple(int(v) for v in p) for p in paddings]):
        flat += [lo, hi]
    if x.dtype == torch.bool:
        return F.pad(x.to(torch.uint8), flat, value=int(bool(constant_values))).bool()
    return F.pad(x, flat, value=constant_values)


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (int(axis),)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    return x.max() if axis is None else x.amax(dim=_axes(axis), keepdim=keepdims)


def reduce_all(x, axis=None):
    return x.all() if axis is None else x.all(dim=axis)


def reduce_any(x, axis=None):
    return x.any() if axis is None else x.any(dim=axis)


def matmul(a, b, transpose_a=False, transpose_b=False):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def einsum(eq, *ops):
    return torch.einsum(eq, *ops)


def tensordot(a, b, axes):
    return torch.tensordot(a, b, dims=axes)


def _num(fn_t, fn_py):
    def f(x, *a):
        if _is_t(x) or any(_is_t(v) for v in a):
            return fn_t(_t(x), *[_t(v) for v in a])
        return fn_py(x, *a)
    return f


def sqrt(x):
    return torch.sqrt(x) if _is_t(x) else torch.tensor(math.sqrt(x), dtype=FLOATX)


def square(x):
    return x * x


exp = _num(torch.exp, math.exp)
log = _num(torch.log, math.log)
abs_ = _num(torch.abs, abs)
erf = _num(torch.erf, math.erf)
tanh = _num(torch.tanh, math.tanh)


def ceil(x):
    return torch.ceil(x) if _is_t(x) else float(math.ceil(x))


def minimum(a, b):
    if _is_t(a) or _is_t(b):
        return torch.minimum(_t(a), _t(b))
    return min(a, b)


def maximum(a, b):
    if _is_t(a) or _is_t(b):
        a, b = _t(a), _t(b)
        if a.dtype != b.dtype:
            a, b = a.to(torch.promote_types(a.dtype, b.dtype)), b.to(torch.promote_types(a.dtype, b.dtype))
        return torch.maximum(a, b)
    return max(a, b)


def equal(a, b):
    if _is_t(a) or _is_t(b):
        return torch.eq(_t(a), _t(b))
    return a == b


def logical_or(a, b):
    return torch.logical_or(_t(a), _t(b)) if (_is_t(a) or _is_t(b)) else (a or b)


def logical_and(a, b):
    return torch.logical_and(_t(a), _t(b)) if (_is_t(a) or _is_t(b)) else (a and b)


def logical_not(a):
    return torch.logical_not(a) if _is_t(a) else (not a)


def where(cond, x=None, y=None):
    if x is None:
        return torch.nonzero(cond)
    x, y = _t(x), _t(y)
    if x.dtype != y.dtype:
        dt_ = torch.promote_types(x.dtype, y.dtype)
        x, y = x.to(dt_), y.to(dt_)
    return torch.where(_t(cond), x, y)


def is_nan(x):
    return torch.isnan(_t(x))


def is_inf(x):
    return torch.isinf(_t(x))


def clip_by_value(x, lo, hi):
    return torch.clamp(x, lo, hi)


def argmin(x, axis=None, output_type=int64):
    """tf.argmin: int64, FIRST minimum on ties (torch.argmin documents the same)."""
    return torch.argmin(x, dim=axis).to(_dt(output_type))


def argmax(x, axis=None, output_type=int64):
    return torch.argmax(x, dim=axis).to(_dt(output_type))


def argsort(x, axis=-1, direction="ASCENDING", stable=False):
    return torch.sort(x, dim=axis, descending=(direction == "DESCENDING"), stable=True).indices.to(torch.int32)


def one_hot(indices, depth, dtype=float32, **kw):
    return F.one_hot(indices.long(), int(depth)).to(_dt(dtype))


def gather(params, indices, axis=0, batch_dims=0):
    """tf.gather. batch_dims=0: take along `axis`; batch_dims=1, axis=1: out[b, ...idx] = params[b, indices[b, ...]]."""
    idx = indices.long()
    if batch_dims == 0:
        out = torch.index_select(params, axis, idx.reshape(-1))
        return out.reshape(params.shape[:axis] + idx.shape + params.shape[axis + 1:])
    assert batch_dims == 1 and axis == 1, "tf_shim.gather: only batch_dims=1 with axis=1"
    B = params.shape[0]
    rows = torch.arange(B).reshape((B,) + (1,) * (idx.dim() - 1)).expand_as(idx)
    return params[rows, idx]


def roll(x, shift, axis):
    return torch.roll(x, int(shift), int(axis))


def tensor_scatter_nd_update(tensor, indices, updates):
    out = tensor.clone()
    idx = indices.long()
    out[tuple(idx[..., i] for i in range(idx.shape[-1]))] = updates
    return out


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def while_loop(cond_fn, body, loop_vars, **kw):
    vars_ = list(loop_vars)
    while bool(cond_fn(*vars_)):
        vars_ = list(body(*vars_))
    return vars_


class TensorArray:
    def __init__(self, dtype, size=0, dynamic_size=False, **kw):
        self._items = [None] * int(size)

    def write(self, i, v):
        i = int(i)
        if i >= len(self._items):
            self._items += [None] * (i + 1 - len(self._items))
        self._items[i] = v
        return self

    def stack(self):
        return torch.stack(self._items, 0)


def timestamp():
    import time
    return torch.tensor(time.time(), dtype=torch.float64)


def clip_by_global_norm(t_list, clip_norm):
    """tf.clip_by_global_norm: global_norm = sqrt(sum_i ||t_i||^2); t_i * clip_norm / max(global_norm, clip_norm)."""
    t_list = list(t_list)
    gn = torch.sqrt(sum((t.detach() ** 2).sum() for t in t_list if t is not None))
    scale = clip_norm / torch.maximum(gn, torch.tensor(float(clip_norm), dtype=gn.dtype))
    return [None if t is None else t * scale for t in t_list], gn


def clip_by_norm(t, clip_norm):
    n = torch.sqrt((t ** 2).sum())
    return t * (clip_norm / torch.maximum(n, torch.tensor(float(clip_norm), dtype=n.dtype)))


def function(func=None, **kw):
    """@tf.function / @tf.function(input_signature=...): eager execution."""
    if func is None:
        return lambda f: f
    return func


class TensorSpec:
    def __init__(self, shape=None, dtype=float32, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


# ---------------------------------------------------------------------------------------------------------------------
# Variable, GradientTape
# ---------------------------------------------------------------------------------------------------------------------
class Variable(torch.nn.Parameter):
    """tf.Variable: a leaf tensor; autograd plays the tape."""

    def __new__(cls, initial_value=None, trainable=True, dtype=None, name=None, **kw):
        if callable(initial_value):
            initial_value = initial_value()
        data = _t(initial_value, dtype).detach().clone()
        if data.dtype == torch.float32 and FLOATX != torch.float32:
            data = data.to(FLOATX)
        v = torch.Tensor._make_subclass(cls, data, bool(trainable) and data.is_floating_point())
        v._tf_name = name or "Variable"
        v._trainable = bool(trainable)
        return v

    def __init__(self, *a, **kw):
        pass

    @property
    def trainable(self):
        return self._trainable

    def assign(self, value):
        with torch.no_grad():
            self.copy_(_t(value).to(self.dtype))
        return self

    def assign_add(self, value):
        with torch.no_grad():
            self.add_(_t(value).to(self.dtype))
        return self

    def assign_sub(self, value):
        with torch.no_grad():
            self.sub_(_t(value).to(self.dtype))
        return self

    def numpy(self):
        return self.detach().numpy()

    def __deepcopy__(self, memo):
        return self


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, t):
        pass

    def gradient(self, target, sources, **kw):
        """tape.gradient: None for sources the target does not depend on."""
        single = _is_t(sources)
        src = [sources] if single else list(sources)
        if not _is_t(target) or not target.requires_grad:
            out = [None] * len(src)
        else:
            out = list(torch.autograd.grad(target, src, allow_unused=True, retain_graph=True))
        return out[0] if single else out


# ---------------------------------------------------------------------------------------------------------------------
# Keras layer machinery
# ---------------------------------------------------------------------------------------------------------------------
class _CallCtx:
    training = None
    depth = 0


_CTX = _CallCtx()
_UID = {}


def _uid(prefix):
    _UID[prefix] = _UID.get(prefix, 0) + 1
    return prefix if _UID[prefix] == 1 else f"{prefix}_{_UID[prefix] - 1}"


def _snake(name):
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and not name[i - 1].isupper():
            out.append("_")
        out.append(ch.lower())
    return "".join(out)


def glorot_uniform(shp, fan_in, fan_out):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(tuple(shp), generator=_RNG, dtype=torch.float64) * 2 - 1) * limit).to(FLOATX)


class Layer:
    """tf.keras.layers.Layer (Keras 2.10 behaviours the reference depends on):
    * weights are created lazily by build(input_shape) on the first __call__;
    * trainable_variables = own weights (creation order) followed by those of tracked sub-layers in attribute-assignment order
      (lists/tuples/dicts of layers are tracked like Keras' ListWrapper), de-duplicated;
    * `training`: an explicit non-None argument wins; otherwise the value of the enclosing layer call is used; otherwise the
      call signature's default. The resolved value is injected if call() accepts `training`, and becomes the context for
      nested calls (so Dropout inside project_hid, called without training=, IS active during a training=True model call)."""

    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        object.__setattr__(self, "_tracked", [])
        object.__setattr__(self, "_own", [])
        self.built = False
        self.trainable = trainable
        self.name = name or _uid(_snake(type(self).__name__))
        self._call_sig = None

    # -- tracking ------------------------------------------------------------------------------------------------
    def __setattr__(self, k, v):
        if not hasattr(self, "_tracked"):
            raise RuntimeError("It looks like you are subclassing `Layer` and forgot to call `super().__init__()` first")
        if isinstance(v, Variable):
            if not any(o is v for o in self._own):
                self._own.append(v)
        elif isinstance(v, Layer):
            if not any(o is v for o in self._tracked):
                self._tracked.append(v)
        elif isinstance(v, (list, tuple)):
            for e in v:
                if isinstance(e, Layer) and not any(o is e for o in self._tracked):
                    self._tracked.append(e)
                elif isinstance(e, Variable) and not any(o is e for o in self._own):
                    self._own.append(e)
            if isinstance(v, list):
                v = _TrackedList(self, v)
        elif isinstance(v, dict):
            for e in v.values():
                if isinstance(e, Layer) and not any(o is e for o in self._tracked):
                    self._tracked.append(e)
        object.__setattr__(self, k, v)

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kw):
        shp = tuple(int(s) for s in (shape or ()))
        if initializer in (None, "glorot_uniform"):
            fan_in, fan_out = (shp[0], shp[-1]) if len(shp) >= 2 else (shp[0] if shp else 1, shp[0] if shp else 1)
            init = glorot_uniform(shp, fan_in, fan_out)
        elif initializer == "zeros":
            init = torch.zeros(shp, dtype=FLOATX)
        elif initializer == "ones":
            init = torch.ones(shp, dtype=FLOATX)
        elif callable(initializer):
            init = initializer(shp)
        else:
            raise NotImplementedError(f"tf_shim: initializer {initializer!r}")
        v = Variable(init, trainable=trainable, name=f"{self.name}/{name}:0")
        self._own.append(v)
        return v

    @property
    def trainable_variables(self):
        out, seen = [], set()

        def visit(layer):
            for v in layer._own:
                if v.trainable and layer.trainable and id(v) not in seen:
                    seen.add(id(v))
                    out.append(v)
            for sub in layer._tracked:
                visit(sub)
        visit(self)
        return out

    trainable_weights = trainable_variables

    @property
    def variables(self):
        return self.trainable_variables

    weights = variables

    def count_params(self):
        return sum(v.numel() for v in self.trainable_variables)

    def build(self, input_shape):
        pass

    def call(self, inputs, *args, **kwargs):
        return inputs

    def __call__(self, *args, **kwargs):
        if self._call_sig is None:
            object.__setattr__(self, "_call_sig", inspect.signature(self.call))
        sig = self._call_sig
        accepts = "training" in sig.parameters
        explicit = None
        if "training" in kwargs:
            explicit = kwargs["training"]
        elif accepts:
            try:
                b = sig.bind_partial(*args, **kwargs)
                if "training" in b.arguments:
                    explicit = b.arguments["training"]
            except TypeError:
                pass
        if explicit is not None:
            training = explicit
        elif _CTX.training is not None:
            training = _CTX.training
        elif accepts and sig.parameters["training"].default not in (inspect._empty, None):
            training = sig.parameters["training"].default
        else:
            training = None if not accepts else False
        if accepts and "training" not in kwargs:
            try:
                b = sig.bind_partial(*args, **kwargs)
                if "training" not in b.arguments:
                    kwargs["training"] = training
                elif b.arguments["training"] is None:
                    args = list(args)
                    pos = list(sig.parameters).index("training")
                    if pos < len(args):
                        args[pos] = training
                    else:
                        kwargs["training"] = training
            except TypeError:
                kwargs["training"] = training
        elif accepts and kwargs.get("training") is None:
            kwargs["training"] = training
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            ishape = tuple(first.shape) if _is_t(first) else None
            self.build(ishape)
            self.built = True
        prev = _CTX.training
        if training is not None:
            _CTX.training = training
        try:
            return self.call(*args, **kwargs)
        finally:
            _CTX.training = prev

    # Keras Model conveniences the scripts touch outside the step
    def compile(self, *a, **k):
        pass

    def summary(self, *a, **k):
        print(f"Model: {self.name}: {self.count_params():,} trainable parameters")

    def save_weights(self, path, *a, **k):
        import os
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save([v.detach().clone() for v in self.trainable_variables], path + ".shim.pt")

    def get_weights(self):
        return [v.detach().numpy() for v in self.trainable_variables]


class _TrackedList(list):
    """Keras wraps list attributes so that later appends are tracked too (self.conv_layers = []; .append(layer))."""

    def __init__(self, owner, items):
        super().__init__(items)
        self._owner = owner

    def append(self, e):
        super().append(e)
        if isinstance(e, Layer) and not any(o is e for o in self._owner._tracked):
            self._owner._tracked.append(e)


class Model(Layer):
    @property
    def layers(self):
        return list(self._tracked)


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._seq = []
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        self._seq.append(layer)
        if not any(o is layer for o in self._tracked):
            self._tracked.append(layer)

    def call(self, inputs, training=None):
        x = inputs
        for l in self._seq:
            x = l(x)          # the call context carries `training`
        return x


class Dense(Layer):
    """y = x @ kernel[in,out] + bias; glorot_uniform / zeros."""

    def __init__(self, units, activation=None, use_bias=True, name=None, **kw):
        super().__init__(name=name or _uid("dense"))
        self.units, self.use_bias = int(units), use_bias
        self.activation = activations_get(activation)

    def build(self, input_shape):
        fin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (fin, self.units), initializer="glorot_uniform")
        self.bias = self.add_weight("bias", (self.units,), initializer="zeros") if self.use_bias else None

    def call(self, x):
        y = x @ self.kernel
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y) if self.activation else y


class Conv1D(Layer):
    """Keras Conv1D, channels-last. kernel [k, Cin/groups, Cout]. padding='same' follows TF's rule:
    out = ceil(T/stride); total = max((out-1)*stride + k - T, 0); left = total // 2, the odd element goes RIGHT.
    groups: input channels split contiguously, group g feeds output channels [g*Cout/G, (g+1)*Cout/G)."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, groups=1, activation=None, name=None, **kw):
        super().__init__(name=name or _uid("conv1d"))
        self.filters = int(filters)
        self.k = int(kernel_size[0] if isinstance(kernel_size, (list, tuple)) else kernel_size)
        self.s = int(strides[0] if isinstance(strides, (list, tuple)) else strides)
        self.padding, self.use_bias, self.groups = padding.lower(), use_bias, int(groups)
        self.activation = activations_get(activation)

    def build(self, input_shape):
        cin = int(input_shape[-1])
        cg = cin // self.groups
        limit_in, limit_out = self.k * cg, self.k * self.filters // self.groups
        self.kernel = self.add_weight("kernel", (self.k, cg, self.filters), initializer=lambda s: glorot_uniform(s, limit_in, limit_out))
        self.bias = self.add_weight("bias", (self.filters,), initializer="zeros") if self.use_bias else None

    def call(self, x):
        T = x.shape[1]
        xt = x.transpose(1, 2)
        if self.padding == "same":
            out = -(-T // self.s)
            total = max((out - 1) * self.s + self.k - T, 0)
            xt = F.pad(xt, (total // 2, total - total // 2))
        y = F.conv1d(xt, self.kernel.permute(2, 1, 0), self.bias, stride=self.s, groups=self.groups).transpose(1, 2)
        return self.activation(y) if self.activation else y


class LayerNormalization(Layer):
    """last axis; biased variance; (x - mean) * rsqrt(var + eps) * gamma + beta."""

    def __init__(self, axis=-1, epsilon=1e-3, name=None, **kw):
        super().__init__(name=name or _uid("layer_normalization"))
        self.epsilon = float(epsilon)

    def build(self, input_shape):
        d = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (d,), initializer="ones")
        self.beta = self.add_weight("beta", (d,), initializer="zeros")

    def call(self, x):
        mean = x.mean(-1, keepdim=True)
        var = ((x - mean) ** 2).mean(-1, keepdim=True)
        return (x - mean) / torch.sqrt(var + self.epsilon) * self.gamma + self.beta


class Dropout(Layer):
    """Inverted dropout when training: keep with probability 1-rate, scale kept values by 1/(1-rate)."""

    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name or _uid("dropout"))
        self.rate = float(rate)
        self.calls_training = 0

    def call(self, x, training=None):
        if not training or self.rate <= 0.0:
            return x
        self.calls_training += 1
        keep = (torch.rand(x.shape, generator=_RNG) >= self.rate).to(x.dtype)
        return x * keep / (1.0 - self.rate)


class Embedding(Layer):
    """uniform(-0.05, 0.05) table; lookup."""

    def __init__(self, input_dim, output_dim, name=None, **kw):
        super().__init__(name=name or _uid("embedding"))
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)

    def build(self, input_shape):
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim),
                                          initializer=lambda s: ((torch.rand(s, generator=_RNG, dtype=torch.float64) * 0.1) - 0.05).to(FLOATX))

    def call(self, ids):
        return self.embeddings[ids.long()]


class Activation(Layer):
    def __init__(self, activation, name=None, **kw):
        super().__init__(name=name or _uid("activation"))
        self.fn = activations_get(activation)

    def call(self, x):
        return self.fn(x)


def gelu(x, approximate=False):
    if approximate:
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def activations_get(a):
    if a is None or callable(a):
        return a
    return {"relu": torch.relu, "gelu": gelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "linear": None,
            "softmax": lambda x: torch.softmax(x, -1)}[a]


# ---- losses / metrics ------------------------------------------------------------------------------------------------
def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None, name=None):
    lse = torch.logsumexp(logits, dim=-1)
    return lse - torch.gather(logits, -1, labels.long().unsqueeze(-1)).squeeze(-1)


def sparse_categorical_crossentropy(y_true, y_pred, from_logits=False, axis=-1):
    if from_logits:
        return sparse_softmax_cross_entropy_with_logits(labels=y_true, logits=y_pred)
    p = torch.gather(y_pred, -1, y_true.long().unsqueeze(-1)).squeeze(-1)
    return -torch.log(torch.clamp(p, 1e-7, 1.0))


class SparseCategoricalCrossentropy:
    def __init__(self, from_logits=False, reduction="auto", name=None):
        self.from_logits, self.reduction = from_logits, reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        l = sparse_categorical_crossentropy(y_true, y_pred, from_logits=self.from_logits)
        return l if self.reduction == "none" else l.mean()


class MeanSquaredError:
    def __init__(self, reduction="auto", name=None):
        self.reduction = reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        l = ((y_true - y_pred) ** 2).mean(-1)
        return l if self.reduction == "none" else l.mean()


class _Metric:
    def __init__(self, name=None, **kw):
        self.name, self._sum, self._n = name, 0.0, 0

    def update_state(self, *a, **k):
        if a and _is_t(a[0]) and a[0].numel() == 1:
            self._sum += float(a[0]); self._n += 1

    def result(self):
        return torch.tensor(self._sum / max(self._n, 1))

    def reset_states(self):
        self._sum, self._n = 0.0, 0

    reset_state = reset_states
    __call__ = update_state


# ---- optimizer ---------------------------------------------------------------------------------------------------------
class _ReplicaCtx:
    strategy = None
    collecting = None      # list that receives (grads, vars) of the running replica when a multi-replica strategy.run is active


class Adam:
    """tf.keras.optimizers.Adam as Keras 2.10 ships it (the legacy OptimizerV2): t = iterations + 1;
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps).
    clipnorm: per-variable g * clipnorm / max(||g||, clipnorm). Inside strategy.run with N replicas, apply_gradients first
    SUMS the replicas' gradients (cross-replica all-reduce, no division) and only then clips and updates — Keras'
    _aggregate_gradients runs before _transform_gradients. Variables whose gradient is None are skipped."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None, clipvalue=None, name="Adam", **kw):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = float(learning_rate), beta_1, beta_2, float(epsilon)
        self.clipnorm = clipnorm
        self.iterations = 0
        self._m, self._v = {}, {}

    lr = property(lambda self: self.learning_rate)

    def apply_gradients(self, grads_and_vars, **kw):
        pairs = [(g, v) for g, v in grads_and_vars]
        if _ReplicaCtx.collecting is not None:
            _ReplicaCtx.collecting.append((self, pairs))
            return
        self._apply(pairs)

    def _apply(self, pairs):
        self.iterations += 1
        t = self.iterations
        lr_t = self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        with torch.no_grad():
            for g, var in pairs:
                if g is None:
                    continue
                g = g.detach()
                if self.clipnorm is not None:
                    g = clip_by_norm(g, self.clipnorm)
                m = self._m.setdefault(id(var), torch.zeros_like(var))
                v = self._v.setdefault(id(var), torch.zeros_like(var))
                m.add_((g - m) * (1.0 - self.beta_1))
                v.add_((g * g - v) * (1.0 - self.beta_2))
                var.sub_(lr_t * m / (torch.sqrt(v) + self.epsilon))


# ---- tf.distribute ---------------------------------------------------------------------------------------------------
class PerReplica:
    def __init__(self, values):
        self.values = list(values)


class MultiWorkerMirroredStrategy:
    """One process. num_replicas (default 1) replicas are emulated sequentially on mirrored (= shared) variables:
    strategy.run(fn, args) calls fn once per replica with that replica's slice of every PerReplica argument; optimizer updates
    issued inside are deferred, their gradients summed across replicas, then applied once (see Adam)."""

    def __init__(self, communication_options=None, cluster_resolver=None, num_replicas=1):
        self.num_replicas_in_sync = int(num_replicas)

    @contextlib.contextmanager
    def scope(self):
        yield

    def experimental_distribute_dataset(self, ds):
        return ds

    def run(self, fn, args=(), kwargs=None):
        kwargs = kwargs or {}
        N = self.num_replicas_in_sync

        def pick(a, r):
            if isinstance(a, PerReplica):
                return a.values[r]
            if isinstance(a, (tuple, list)) and any(isinstance(e, PerReplica) for e in a):
                return type(a)(pick(e, r) for e in a)
            return a
        if N == 1:
            return fn(*[pick(a, 0) for a in args], **kwargs)
        outs, collected = [], []
        for r in range(N):
            _ReplicaCtx.collecting = []
            try:
                outs.append(fn(*[pick(a, r) for a in args], **kwargs))
            finally:
                collected.append(_ReplicaCtx.collecting)
                _ReplicaCtx.collecting = None
        n_apply = len(collected[0])
        for i in range(n_apply):
            opt, pairs0 = collected[0][i]
            summed = []
            for j, (g0, var) in enumerate(pairs0):
                gs = [collected[r][i][1][j][0] for r in range(N)]
                summed.append((None if all(g is None for g in gs) else sum(g.detach() for g in gs if g is not None), var))
            opt._apply(summed)
        return PerReplica(outs)

    def reduce(self, reduce_op, value, axis=None):
        if isinstance(value, PerReplica):
            tot = sum(_t(v).detach() for v in value.values)
            return tot / len(value.values) if str(reduce_op).upper().endswith("MEAN") else tot
        return value


class _Dataset:
    def __init__(self, make_iter):
        self._make = make_iter

    @staticmethod
    def from_tensor_slices(tensors):
        def gen():
            if isinstance(tensors, (tuple, list)):
                ts = [_t(t) for t in tensors]
                for i in range(ts[0].shape[0]):
                    yield tuple(t[i] for t in ts)
            else:
                t = _t(tensors)
                for i in range(t.shape[0]):
                    yield t[i]
        return _Dataset(gen)

    @staticmethod
    def from_generator(generator, output_types=None, output_shapes=None, output_signature=None, args=None):
        def gen():
            for item in generator():
                if isinstance(item, tuple):
                    if output_signature is not None:
                        yield tuple(_t(e, s.dtype) for e, s in zip(item, output_signature))
                    else:
                        yield tuple(_t(e) for e in item)
                else:
                    yield _t(item)
        return _Dataset(gen)

    def batch(self, n, drop_remainder=False, **kw):
        src = self._make

        def gen():
            buf = []
            for item in src():
                buf.append(item)
                if len(buf) == n:
                    yield _collate(buf)
                    buf = []
            if buf and not drop_remainder:
                yield _collate(buf)
        return _Dataset(gen)

    def repeat(self, count=None):
        src = self._make

        def gen():
            k = 0
            while count is None or k < count:
                empty = True
                for item in src():
                    empty = False
                    yield item
                if empty:
                    return
                k += 1
        return _Dataset(gen)

    def prefetch(self, *a, **k):
        return self

    def cache(self, *a, **k):
        return self

    def shuffle(self, *a, **k):
        return self

    def map(self, fn, **k):
        src = self._make
        return _Dataset(lambda: (fn(*i) if isinstance(i, tuple) else fn(i) for i in src()))

    def take(self, n):
        src = self._make

        def gen():
            for i, item in enumerate(src()):
                if i >= n:
                    return
                yield item
        return _Dataset(gen)

    def __iter__(self):
        return iter(self._make())


def _collate(buf):
    if isinstance(buf[0], tuple):
        return tuple(torch.stack([b[i] for b in buf]) for i in range(len(buf[0])))
    return torch.stack(buf)


class _Checkpoint:
    def __init__(self, **kw):
        self.saved = []

    def save(self, file_prefix=None, **kw):
        self.saved.append(file_prefix)
        return f"{file_prefix}-{len(self.saved)}"

    def restore(self, *a, **k):
        return self


# ---- tf.random / tf.nn / tf.linalg / tf.signal --------------------------------------------------------------------------
def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
    t = (torch.randn(tuple(int(s) for s in shape), generator=_RNG, dtype=torch.float64) * stddev + mean).to(_dt(dtype))
    RANDOM_LOG.append(("normal", t))
    return t


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None):
    dt_ = _dt(dtype)
    shp = tuple(int(s) for s in shape)
    if dt_.is_floating_point:
        hi = 1.0 if maxval is None else float(maxval)
        t = (torch.rand(shp, generator=_RNG, dtype=torch.float64) * (hi - minval) + minval).to(dt_)
    else:
        t = torch.randint(int(minval), int(maxval), shp, generator=_RNG).to(dt_)
    RANDOM_LOG.append(("uniform", t))
    return t


def random_shuffle(value, seed=None):
    perm = torch.randperm(value.shape[0], generator=_RNG)
    t = value[perm]
    RANDOM_LOG.append(("shuffle", t))
    return t


def softmax(x, axis=-1):
    return torch.softmax(x, dim=axis)


def top_k(x, k=1, sorted=True):
    """tf.nn.top_k: values descending; among equal values the LOWER index comes first (a stable descending sort)."""
    r = torch.sort(x, dim=-1, descending=True, stable=True)
    return r.values[..., :int(k)], r.indices[..., :int(k)].to(torch.int32)


def moments(x, axes, keepdims=False):
    mean = x.mean(dim=tuple(axes), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=tuple(axes), keepdim=True)
    if not keepdims:
        mean, var = mean.squeeze(tuple(axes)), var.squeeze(tuple(axes))
    return mean, var


def band_part(x, num_lower, num_upper):
    """keep element (i, j) iff (num_lower < 0 or i - j <= num_lower) and (num_upper < 0 or j - i <= num_upper)."""
    m, n = x.shape[-2], x.shape[-1]
    i = torch.arange(m).unsqueeze(1)
    j = torch.arange(n).unsqueeze(0)
    keep = torch.ones(m, n, dtype=torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return x * keep.to(x.dtype)


def ctc_loss(labels, logits, label_length, logit_length, logits_time_major=True, blank_index=None, **kw):
    """tf.nn.ctc_loss (dense labels): per-example negative log-likelihood; blank_index defaults to 0 for dense labels;
    logits are unnormalised (log_softmax applied inside). torch's F.ctc_loss implements the same forward-backward
    recursion and serves here as an implementation independent of oracle/."""
    lg = logits if logits_time_major else logits.transpose(0, 1)         # [T, B, C]
    blank = 0 if blank_index is None else int(blank_index)
    if blank < 0:
        blank += lg.shape[-1]
    lp = torch.log_softmax(lg, dim=-1)
    return F.ctc_loss(lp, labels.long(), _t(logit_length).long(), _t(label_length).long(), blank=blank, reduction="none",
                      zero_infinity=False)


def hann_window(window_length, periodic=True, dtype=float32):
    n = torch.arange(int(window_length), dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / (window_length if periodic else window_length - 1))).to(_dt(dtype))


def stft(signals, frame_length, frame_step, fft_length=None, window_fn=hann_window, pad_end=False, name=None):
    """tf.signal.stft: frames of frame_length every frame_step (no padding), periodic Hann window, rfft of fft_length."""
    fft_length = fft_length or int(2 ** math.ceil(math.log2(frame_length)))
    frames = signals.unfold(-1, frame_length, frame_step)
    if window_fn is not None:
        frames = frames * window_fn(frame_length, dtype=frames.dtype)
    return torch.fft.rfft(frames, n=fft_length, dim=-1)


def linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000, lower_edge_hertz=125.0,
                                upper_edge_hertz=3800.0, dtype=float32, name=None):
    """tf.signal.linear_to_mel_weight_matrix: HTK mel = 1127 ln(1 + f/700); the DC bin is dropped (zero row); band edges are
    num_mel_bins + 2 points linear in mel; triangles min(lower slope, upper slope) clipped at 0, computed in mel space."""
    def mel(f):
        return 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)
    lin = np.linspace(0.0, sample_rate / 2.0, num_spectrogram_bins)[1:]
    sm = mel(lin)[:, None]
    edges = np.linspace(mel(lower_edge_hertz), mel(upper_edge_hertz), num_mel_bins + 2)
    lo, ce, up = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
    w = np.maximum(0.0, np.minimum((sm - lo) / (ce - lo), (up - sm) / (up - ce)))
    return torch.as_tensor(np.pad(w, [[1, 0], [0, 0]])).to(_dt(dtype))


# ---------------------------------------------------------------------------------------------------------------------
# module assembly
# ---------------------------------------------------------------------------------------------------------------------
def _ns(name, **kw):
    m = types.ModuleType(name)
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def build_module():
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "2.10.0-shim"
    g = globals()
    for name in ("shape constant convert_to_tensor cast zeros ones fill zeros_like ones_like reshape transpose expand_dims squeeze "
                 "concat stack tile pad reduce_sum reduce_mean reduce_max reduce_all reduce_any matmul einsum tensordot sqrt square exp "
                 "minimum maximum equal logical_or logical_and logical_not where clip_by_value argmin argmax argsort one_hot gather "
                 "roll tensor_scatter_nd_update cond while_loop TensorArray timestamp clip_by_global_norm clip_by_norm function "
                 "TensorSpec Variable GradientTape float32 float64 float16 int32 int64 tanh").split():
        setattr(tf, name, g[name])
    tf.range = range_
    tf.abs = abs_
    tf.bool = bool_
    tf.Tensor = torch.Tensor
    tf.math = _ns("tensorflow.math", sqrt=sqrt, erf=erf, log=log, exp=exp, ceil=ceil, is_nan=is_nan, is_inf=is_inf, square=square,
                  abs=abs_, reduce_sum=reduce_sum, reduce_mean=reduce_mean, maximum=maximum, minimum=minimum, tanh=tanh,
                  argmax=argmax, argmin=argmin, top_k=top_k, softmax=softmax, equal=equal, logical_not=logical_not)
    tf.nn = _ns("tensorflow.nn", softmax=softmax, moments=moments, top_k=top_k, ctc_loss=ctc_loss, gelu=gelu, relu=torch.relu,
                sparse_softmax_cross_entropy_with_logits=sparse_softmax_cross_entropy_with_logits,
                log_softmax=lambda x, axis=-1: torch.log_softmax(x, dim=axis))
    tf.linalg = _ns("tensorflow.linalg", band_part=band_part, matmul=matmul)
    tf.random = _ns("tensorflow.random", normal=random_normal, uniform=random_uniform, shuffle=random_shuffle, set_seed=seed)
    tf.signal = _ns("tensorflow.signal", stft=stft, hann_window=hann_window, linear_to_mel_weight_matrix=linear_to_mel_weight_matrix)
    layers = _ns("tensorflow.keras.layers", Layer=Layer, Dense=Dense, Conv1D=Conv1D, LayerNormalization=LayerNormalization,
                 Dropout=Dropout, Embedding=Embedding, Activation=Activation)
    losses = _ns("tensorflow.keras.losses", SparseCategoricalCrossentropy=SparseCategoricalCrossentropy,
                 sparse_categorical_crossentropy=sparse_categorical_crossentropy, MeanSquaredError=MeanSquaredError,
                 Reduction=_ns("Reduction", NONE="none", SUM="sum", AUTO="auto", SUM_OVER_BATCH_SIZE="sum_over_batch_size"))
    metrics = _ns("tensorflow.keras.metrics", Mean=_Metric, SparseCategoricalAccuracy=_Metric)
    tf.keras = _ns("tensorflow.keras", layers=layers, losses=losses, metrics=metrics, Model=Model, Sequential=Sequential,
                   optimizers=_ns("tensorflow.keras.optimizers", Adam=Adam),
                   activations=_ns("tensorflow.keras.activations", gelu=gelu, get=activations_get, relu=torch.relu, tanh=torch.tanh))
    tf.train = _ns("tensorflow.train", Checkpoint=_Checkpoint)
    exp_ = _ns("tensorflow.distribute.experimental", CommunicationOptions=lambda **k: dict(k),
               CommunicationImplementation=_ns("CommunicationImplementation", NCCL="NCCL", AUTO="AUTO", RING="RING"),
               MultiWorkerMirroredStrategy=MultiWorkerMirroredStrategy)
    tf.distribute = _ns("tensorflow.distribute", MultiWorkerMirroredStrategy=MultiWorkerMirroredStrategy, experimental=exp_,
                        ReduceOp=_ns("ReduceOp", SUM="SUM", MEAN="MEAN"), PerReplica=PerReplica)
    tf.data = _ns("tensorflow.data", Dataset=_Dataset, AUTOTUNE=-1, experimental=_ns("tensorflow.data.experimental", AUTOTUNE=-1))
    tf.config = _ns("tensorflow.config", experimental=_ns("tensorflow.config.experimental", list_physical_devices=lambda *a: [],
                                                           set_memory_growth=lambda *a: None),
                    list_physical_devices=lambda *a: [])
    tf._shim = sys.modules[__name__]
    return tf


def install():
    """Put the stand-in into sys.modules['tensorflow'] (+ the sub-modules the scripts import with `from tensorflow.keras
    import layers, Model`); refuses to shadow a real TensorFlow."""
    cur = sys.modules.get("tensorflow")
    if cur is not None and not hasattr(cur, "_shim"):
        raise RuntimeError("a real tensorflow is importable: use it instead of the shim")
    if cur is None:
        tf = build_module()
        tf.__path__ = []                                   # lets `import tensorflow.keras` resolve through sys.modules
        sys.modules["tensorflow"] = tf

        def reg(mod):
            sys.modules[mod.__name__] = mod
            for v in vars(mod).values():
                if isinstance(v, types.ModuleType) and v.__name__.startswith("tensorflow.") and v.__name__ not in sys.modules:
                    reg(v)
        for v in list(vars(tf).values()):
            if isinstance(v, types.ModuleType) and v.__name__.startswith("tensorflow."):
                reg(v)
    return sys.modules["tensorflow"]
