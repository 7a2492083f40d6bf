"""TEST INFRASTRUCTURE ONLY — the checkpoint keys tf.train.Checkpoint(model=model, optimizer=optimizer) would give the variables
of a reference model object (running on oracle/tf_shim.py), restated from TensorFlow 2.10's object-graph naming rules:

  * the graph is walked breadth first from the root Checkpoint object; a variable's key is the FIRST path that reaches it
    (tensorflow/python/checkpoint/graph_view.py, util.py `_serialize_object_graph`: `_breadth_first_traversal`);
  * an edge is named after the Python attribute a trackable was assigned to (AutoTrackable.__setattr__), in assignment order;
    variables created by Layer.add_weight(name=...) hang under that `name` (base_layer.add_weight -> _add_variable_with_custom_getter);
  * list attributes are ListWrappers whose children are named by index (`layers/0`);
  * a tf.keras.Sequential (Functional) puts its layers FIRST, as `layer_with_weights-<k>` (k counts layers that own weights) and
    `layer-<i>` (keras/engine/functional.py `_layer_checkpoint_dependencies`);
  * the key is "/".join(edge names) + "/.ATTRIBUTES/VARIABLE_VALUE" (graph_view `_escape_local_name`; none of the reference's
    names needs escaping).

TensorFlow itself cannot be imported here (SURVEY §8 c), so these rules are a restatement, not an execution of TF; what IS
executed is the reference's own constructors (their attribute structure). Used by tests/test_checkpoint.py to pin
tethys_speech_b200.checkpoint.tf_object_key."""
from collections import OrderedDict, deque

from . import tf_shim as S

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def _children(obj):
    """(edge name, child) pairs of one trackable, in TF's dependency order."""
    out = []
    if isinstance(obj, (list, tuple)):
        return [(str(i), e) for i, e in enumerate(obj) if isinstance(e, (S.Layer, S.Variable, list, tuple, dict))]
    if isinstance(obj, dict):
        return [(str(k), e) for k, e in obj.items() if isinstance(e, (S.Layer, S.Variable, list, tuple, dict))]
    if isinstance(obj, S.Sequential):
        k = 0
        for i, layer in enumerate(obj._seq):
            if layer.trainable_variables:
                out.append((f"layer_with_weights-{k}", layer))
                k += 1
            out.append((f"layer-{i}", layer))
    for name, v in vars(obj).items():
        if name in ("_tracked", "_own", "_seq", "_call_sig", "_owner"):       # shim bookkeeping, not reference attributes
            continue
        if isinstance(v, (S.Layer, S.Variable)):
            out.append((name, v))
        elif isinstance(v, (list, tuple)) and any(isinstance(e, (S.Layer, S.Variable)) for e in v):
            out.append((name, v))
        elif isinstance(v, dict) and any(isinstance(e, (S.Layer, S.Variable)) for e in v.values()):
            out.append((name, v))
    return out


def variable_keys(root_children):
    """root_children: e.g. {"model": model}. -> OrderedDict id(variable) -> checkpoint key (first BFS path)."""
    keys = OrderedDict()
    seen = set()
    q = deque()
    for name, obj in root_children.items():
        q.append((name, obj))
    while q:
        path, obj = q.popleft()
        if id(obj) in seen:
            continue
        seen.add(id(obj))
        if isinstance(obj, S.Variable):
            keys[id(obj)] = path + SUFFIX
            continue
        for name, child in _children(obj):
            if id(child) not in seen:
                q.append((path + "/" + name, child))
    return keys
