"""Train loops and CLI plumbing shared by the drop-in scripts under speech_jobs/ — the thin L3/L4 layers of the
reference (W:894-958, W:990-1058; V:1263-1376, V:1380-1487; VS:1180-1292; WS:1183-1306) around the native step.
Same flags, same `Step N, Loss: x.xxxx, Time: …` log line, same jct file; hard-coded /workspace and /result paths are
overridable (TETHYS_WORKSPACE / TETHYS_RESULT) and failures to write them are non-fatal, as SURVEY §5.6 asks."""
import json
import re
import os
import time

import torch

from . import checkpoint
from . import wav2vec2 as W2V
from . import whisper as WH
from .runtime import Adam, GraphedTrainStep, Strategy, to_device

WORKSPACE = os.environ.get("TETHYS_WORKSPACE", "/workspace")
RESULT = os.environ.get("TETHYS_RESULT", "/result")


def task_from_tf_config():
    """W:1037-1040 / job_name.py:3-13: task type and index from TF_CONFIG (None when unset)."""
    tf_config = json.loads(os.environ.get("TF_CONFIG") or "{}")
    task = tf_config.get("task", {})
    return task.get("type"), task.get("index")


def _log_step(step, loss_value, start_time, step_duration):
    elapsed = time.time() - start_time
    print(f"Step {step}, Loss: {loss_value:.4f}, Time: {time.strftime('%H:%M:%S')} (경과: {elapsed:.2f}초, 스텝 시간: {step_duration:.2f}초)", flush=True)


_CHECKPOINTS = {}


def _save_checkpoint(model, name, optimizer=None):
    """checkpoint.save(os.path.join(checkpoint_dir, name)) — V:1341, V:1362, W:956: model + optimizer slots. Goes through one
    checkpoint.Checkpoint object per model, like the reference's tf.train.Checkpoint (V:1286-1288): files are numbered by its
    save counter (`model_step_50-1.tsckpt`, `model_epoch_1-2.tsckpt`), which is what `--resume <dir>` / latest_checkpoint
    orders by."""
    try:
        d = os.path.join(WORKSPACE, "checkpoints")
        os.makedirs(d, exist_ok=True)
        ck = _CHECKPOINTS.get(id(model))
        if ck is None or ck.optimizer is not optimizer:
            ck = _CHECKPOINTS[id(model)] = checkpoint.Checkpoint(model=model, optimizer=optimizer)
            last = checkpoint.latest_checkpoint(d)
            if last:                                  # continue the numbering of an earlier (resumed) run
                m = re.search(r"-(\d+)\.tsckpt$", last)
                ck.save_counter = int(m.group(1)) if m else 0
        return ck.save(os.path.join(d, name))
    except Exception as e:  # the reference's checkpoint dir is container-specific
        print(f"checkpoint not written: {e}")
        return None


def _maybe_resume(resume_from, model, optimizer):
    """Restore model + optimizer from a checkpoint file (or the newest one of a directory); returns the step to continue at.
    Every rank reads the same file, so the replicas stay identical without a second broadcast."""
    if not resume_from:
        return 0
    path = checkpoint.latest_checkpoint(resume_from) if os.path.isdir(resume_from) else resume_from
    if path is None:
        print(f"no checkpoint under {resume_from}; starting from scratch")
        return 0
    meta = checkpoint.restore(path, model, optimizer)
    print(f"restored {path} (iterations = {optimizer.iterations})")
    return int(meta.get("optimizer", {}).get("iterations", 0))


def write_jct(jct, task_type, task_index):
    """W:1013-1021: /result/<model.txt>/<task>_<idx>_jct.txt with '%.2f'."""
    print("jct:", jct)
    try:
        with open(os.path.join(WORKSPACE, "model.txt")) as f:
            save_dir_name = f.read().strip()
        path = os.path.join(RESULT, save_dir_name, f"{task_type}_{task_index}_jct.txt")
        with open(path, "w") as f:
            f.write("%.2f" % float(jct))
    except Exception as e:
        print(f"JCT file not written: {e}")


class _StepRunner:
    """Runs the train step eagerly for the first batches of a given shape (which also allocates the workspace, the optimiser
    state and the gradient buckets), then captures it as CUDA graph(s) and replays them for every later batch of that shape:
    same arithmetic, none of the ~300-430 launch gaps of a step (runtime.GraphedTrainStep / GraphedSegments). A batch of another
    shape (the ragged last batch of an epoch) is run eagerly and the graph is rebuilt afterwards, because the eager call
    re-plans the workspace the captured kernels point into."""

    def __init__(self, eager, build, enabled, eager_steps=2):
        self.eager, self.build, self.enabled, self.eager_steps = eager, build, enabled, eager_steps
        self.graph, self.key, self.count = None, None, 0

    def __call__(self, feats, labels):
        if not self.enabled:
            return self.eager(feats, labels)
        key = (tuple(feats.shape), None if labels is None else tuple(labels.shape))
        if self.graph is not None and key == self.key:
            return self.graph(feats, labels)
        self.graph = None
        self.count = self.count + 1 if key == self.key else 1
        self.key = key
        loss = self.eager(feats, labels)
        if self.count >= self.eager_steps:
            try:
                self.graph = self.build(feats, labels)
            except Exception as e:      # noqa: BLE001 — a failed capture must not end the run: stay eager
                print(f"CUDA-graph capture failed ({e}); continuing with eager steps")
                self.enabled = False
        return loss


def _use_graph(cuda_graph):
    return bool(cuda_graph) and not os.environ.get("TETHYS_NO_CUDA_GRAPH")


def train_whisper(strategy, model_type="small", num_epochs=1, learning_rate=1e-4, batch_size=1, num_batches=40,
                  precision="bf16", seq_len=3000, from_waveform=False, resume_from=None, cuda_graph=True):
    """W:894-958. from_waveform=True (extension, SURVEY f-1): the dataset yields raw 30 s waveforms and the fused log-mel kernel
    (extract_fbank_features, W:739-766) produces the model input inside the loop. resume_from (extension, SURVEY f-3): a
    checkpoint file, or a directory whose newest checkpoint is taken, restored into model + optimizer before the loop."""
    with strategy.scope():
        model = WH.create_whisper_model(model_type=model_type, precision=precision, device=strategy.local_rank)
        model.broadcast_weights(strategy)
        optimizer = Adam(learning_rate=learning_rate)
    step = _maybe_resume(resume_from, model, optimizer)
    global_batch = batch_size * strategy.num_replicas_in_sync
    dataset = (WH.create_dummy_waveform_dataset(global_batch, audio_seconds=seq_len / 100.0) if from_waveform
               else WH.create_dummy_dataset(global_batch, seq_len=seq_len))
    world = strategy.num_replicas_in_sync

    def eager(feats, labels):
        return WH.distributed_train_step(strategy, model, (feats, labels), optimizer)

    def build(feats, labels):
        dev = model._prog.device
        f, lab = to_device(feats, torch.float32, dev), to_device(labels, torch.int32, dev)
        if world > 1:
            gstep, _ = WH.make_graphed_distributed_step(strategy, model, optimizer, f, lab, warmup=0)
            return gstep
        graphed = GraphedTrainStep(lambda batch, aux: WH.train_step(model, batch, optimizer), model, optimizer, (f, lab), None, warmup=0)
        return lambda a, b: graphed((a, b), None)

    run_step = _StepRunner(eager, build, _use_graph(cuda_graph))
    start_time = time.time()
    for epoch in range(num_epochs):
        print(f"Epoch {epoch + 1}/{num_epochs}")
        for _ in range(num_batches):
            feats, labels = next(dataset)
            if world > 1 and feats.shape[0] < global_batch:
                # the ragged last batch of the 50-sample set (W:815) would leave some replicas without samples; every rank skips
                # it so that the all-reduce stays matched (the reference lets TF run empty per-replica batches)
                continue
            lo = strategy.rank * batch_size                      # rank r takes samples [r*B, (r+1)*B) of the global batch
            feats, labels = feats[lo:lo + batch_size], labels[lo:lo + batch_size]
            if feats.shape[0] == 0:
                continue
            step_start = time.time()
            if from_waveform:
                feats = WH.waveform_to_features(feats, device=strategy.local_rank)
            loss = run_step(feats, labels)
            loss_value = float(loss)                              # device sync (the reference's loss.numpy(), W:951)
            _log_step(step, loss_value, start_time, time.time() - step_start)
            step += 1
        if strategy.rank == 0:
            _save_checkpoint(model, f"whisper_{model_type}_epoch_{epoch + 1}", optimizer)
    return model


def train_wav2vec2(strategy, model_type="pretraining", model_size="small", num_epochs=1, learning_rate=3e-5, batch_size=1,
                   num_batches=5, precision="bf16", audio_length=32000, legacy=False, resume_from=None, cuda_graph=True):
    """V:1263-1376 (legacy=True: the whisper_single.py / stable_jobs variant — WS:1183-1258: 5 s audio, unscaled loss,
    no clipping, Adam eps 1e-7, seed-42 shuffle sampler)."""
    with strategy.scope():
        model = W2V.create_full_model(model_type=model_type, model_size=model_size, precision=precision, device=strategy.local_rank)
        model.broadcast_weights(strategy)
        optimizer = Adam(learning_rate=learning_rate, epsilon=1e-7) if legacy else Adam(learning_rate=learning_rate, epsilon=1e-8, clipnorm=1.0)
    step = _maybe_resume(resume_from, model, optimizer)
    global_batch = batch_size * strategy.num_replicas_in_sync
    dataset = W2V.create_dummy_dataset(global_batch, audio_length=audio_length)
    neg_legacy = None
    if legacy:
        T = model.num_frames(audio_length)
        perm = torch.randperm(T, generator=torch.Generator().manual_seed(42))        # tf.random.shuffle(range(T), seed=42) stand-in
        t = torch.arange(T).unsqueeze(1)
        k = torch.arange(model.num_negatives).unsqueeze(0)
        neg_legacy = perm[(k - (t + 1)) % T].to(torch.int32).unsqueeze(0).expand(batch_size, -1, -1).contiguous()   # WS:799-839
    world = strategy.num_replicas_in_sync

    def eager(feats, labels):
        return W2V.distributed_train_step(strategy, model, (feats, labels), optimizer)

    def build(feats, labels):
        dev = model._prog.device
        f = to_device(feats, torch.float32, dev)
        if world > 1:
            gstep, _ = W2V.make_graphed_distributed_step(strategy, model, optimizer, f, warmup=0)
            return lambda a, b: gstep(a)
        T = model.num_frames(f.shape[1])

        def sample_aux():
            return {"neg": model._sample_negative_indices(T, f.shape[0])[:, 0, :].contiguous()}     # V:907-937, outside the graph

        graphed = GraphedTrainStep(lambda batch, aux: W2V.train_step(model, batch, optimizer, neg_indices=aux["neg"]), model, optimizer,
                                   (f, None), sample_aux(), warmup=0)
        return lambda a, b: graphed((a, None), sample_aux())

    # the graph path covers the pre-training step; the task heads and the legacy step stay eager
    run_step = _StepRunner(eager, build, _use_graph(cuda_graph) and model_type == "pretraining" and not legacy)
    start_time = time.time()
    for epoch in range(num_epochs):
        print(f"Epoch {epoch + 1}/{num_epochs}")
        for _ in range(num_batches):
            try:
                feats, labels = next(dataset)
                lo = strategy.rank * batch_size
                feats, labels = feats[lo:lo + batch_size], labels[lo:lo + batch_size]
                step_start = time.time()
                if legacy:
                    loss = W2V.legacy_train_step(model, (feats, None), optimizer, neg_indices=neg_legacy)
                    loss = strategy.reduce("SUM", loss)
                else:
                    loss = run_step(feats, labels)
                try:
                    loss_value = float(loss)
                except Exception:
                    loss_value = 0.0
                _log_step(step, loss_value, start_time, time.time() - step_start)
                step += 1
                if step % 50 == 0 and strategy.rank == 0:
                    _save_checkpoint(model, f"model_step_{step}", optimizer)
            except StopIteration:
                break
            except Exception as e:
                # V:1367-1371 prints this line, restarts the dataset iterator and carries on. Deliberate deviation: the same line is
                # printed, then the error is re-raised (as W:837-842 does) — a failed CUDA launch or collective is sticky, a loop that
                # swallows it would log the same error for every remaining step and leave the other replicas waiting in NCCL.
                print(f"Error at step {step}: {e}")
                raise
        if strategy.rank == 0:
            _save_checkpoint(model, f"model_epoch_{epoch + 1}", optimizer)
    return model


def make_strategy():
    """MultiWorkerMirroredStrategy() — W:1047 / V:1473: under torchrun one process per GPU; alone, one replica."""
    s = Strategy()
    if torch.cuda.is_available():
        torch.cuda.set_device(s.local_rank)
    return s
