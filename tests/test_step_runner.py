"""train._StepRunner — when the CLI train loops switch from eager steps to CUDA-graph replays (host logic only, no GPU):
eager for the first batches of a shape, one capture, replays afterwards; a batch of another shape (the ragged last batch of
an epoch, W:815) runs eagerly and invalidates the graph, which is rebuilt once the regular shape is back; a failed capture
turns the feature off instead of ending the run."""
import torch

from tethys_speech_b200 import train


class _Probe:
    def __init__(self, fail_build=False):
        self.log, self.fail_build = [], fail_build

    def eager(self, feats, labels):
        self.log.append(("eager", tuple(feats.shape)))
        return 1.0

    def build(self, feats, labels):
        if self.fail_build:
            raise RuntimeError("capture not possible")
        self.log.append(("build", tuple(feats.shape)))

        def replay(f, l):
            self.log.append(("graph", tuple(f.shape)))
            return 2.0
        return replay


def test_eager_then_capture_then_replay_and_ragged_batch():
    p = _Probe()
    run = train._StepRunner(p.eager, p.build, enabled=True, eager_steps=2)
    full, ragged = torch.zeros(4, 8), torch.zeros(2, 8)
    lab4, lab2 = torch.zeros(4, 3), torch.zeros(2, 3)
    out = [run(full, lab4) for _ in range(4)]
    assert out == [1.0, 1.0, 2.0, 2.0]
    assert p.log == [("eager", (4, 8)), ("eager", (4, 8)), ("build", (4, 8)), ("graph", (4, 8)), ("graph", (4, 8))]
    p.log.clear()
    assert run(ragged, lab2) == 1.0                         # other shape: eager, graph dropped
    assert [run(full, lab4) for _ in range(3)] == [1.0, 1.0, 2.0]
    assert p.log == [("eager", (2, 8)), ("eager", (4, 8)), ("eager", (4, 8)), ("build", (4, 8)), ("graph", (4, 8))]


def test_labels_shape_is_part_of_the_key_and_none_labels_work():
    p = _Probe()
    run = train._StepRunner(p.eager, p.build, enabled=True, eager_steps=1)
    x = torch.zeros(2, 5)
    assert run(x, None) == 1.0 and run(x, None) == 2.0      # built after the first eager step
    assert run(x, torch.zeros(2, 7)) == 1.0                 # same features, new label shape: not the captured step
    assert p.log[-2:] == [("eager", (2, 5)), ("build", (2, 5))]   # (eager_steps = 1: rebuilt right after that eager step)


def test_disabled_and_failed_capture_stay_eager(capsys):
    p = _Probe()
    run = train._StepRunner(p.eager, p.build, enabled=False)
    assert [run(torch.zeros(1, 2), None) for _ in range(5)] == [1.0] * 5 and all(k == "eager" for k, _ in p.log)
    q = _Probe(fail_build=True)
    run = train._StepRunner(q.eager, q.build, enabled=True, eager_steps=1)
    assert [run(torch.zeros(1, 2), None) for _ in range(4)] == [1.0] * 4
    assert "capture failed" in capsys.readouterr().out and run.enabled is False
    assert sum(k == "build" for k, _ in q.log) == 0


def test_graph_switch_env(monkeypatch):
    assert train._use_graph(True) and not train._use_graph(False)
    monkeypatch.setenv("TETHYS_NO_CUDA_GRAPH", "1")
    assert not train._use_graph(True)
