#!/usr/bin/env python
"""CPU experiment (oracle only, no GPU): how far does a Wav2Vec2 run move when the cross-replica gradient SUM is carried in bf16
buckets instead of fp32 — everything else exact (fp64)? Same setting as tools/check_dist_graph.py's bf16-bucket check: tiny preset,
2 x 6400 samples per replica, dropout off, Adam(1e-4, eps 1e-8, clipnorm 1), local clip_by_global_norm(1) before the reduce.
Two runs per replica count N: (a) exact sum, (b) every replica's clipped gradient rounded to bf16 (ts_grad_pack_bf16) and the sum
rounded to bf16 again (the bucket NCCL returns). As a yardstick, (c) perturbs run (a)'s INITIAL weights by one part in 1e5 — far below
anything bf16 compute does — to show how fast this rapidly-descending toy run amplifies any perturbation.
    python tools/bf16_bucket_drift.py [--steps 4] [--replicas 2 8]  ->  profiles/r02g_bf16_bucket_drift.log"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import wav2vec2_oracle as O


def run(N, steps, mode, seed=0):
    cfg = O.Wav2Vec2Config("tiny")
    w = O.init_weights(cfg, seed=0, dtype=torch.float64)
    if mode == "perturbed":
        g = torch.Generator().manual_seed(99)
        w = {k: v * (1.0 + 1e-5 * torch.randn(v.shape, generator=g, dtype=torch.float64)) for k, v in w.items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v_ = {k: torch.zeros_like(v) for k, v in w.items()}
    T = O.num_frames(cfg, 6400)
    data = []
    for r in range(N):
        g = torch.Generator().manual_seed(100 + r)
        wave = torch.randn(2, 6400, generator=g, dtype=torch.float64)
        neg = O.negative_indices_from_random(torch.randint(0, T, (2, T), generator=g), cfg.num_negatives)
        data.append((wave, neg))
    losses = []
    names = list(w)
    for t in range(1, steps + 1):
        peers, loss = [], 0.0
        for r in range(1, N):
            out, g = O.loss_and_grads(cfg, w, data[r][0], data[r][1], loss_div=float(N))
            gl, _ = O.T.clip_by_global_norm([g[k] for k in names], 1.0)
            if mode == "bf16":
                gl = [x.to(torch.bfloat16).to(torch.float64) for x in gl]
            peers.append(dict(zip(names, gl)))
            loss += float(out["loss"].detach()) / N
        if mode == "bf16":
            # replica 0's own contribution is rounded like the others, and the reduced bucket is bf16 again
            out, g = O.loss_and_grads(cfg, w, data[0][0], data[0][1], loss_div=float(N))
            gl, _ = O.T.clip_by_global_norm([g[k] for k in names], 1.0)
            tot = [x.to(torch.bfloat16).to(torch.float64) for x in gl]
            for pg in peers:
                tot = [(a + pg[k]).to(torch.bfloat16).to(torch.float64) for a, k in zip(tot, names)]
            tot = O.T.clip_by_norm_each(tot, 1.0)
            O.T.keras_adam_step([w[k] for k in names], tot, [m[k] for k in names], [v_[k] for k in names], t, 1e-4, eps=1e-8)
        else:
            out = O.train_step(cfg, w, m, v_, t, data[0][0], data[0][1], lr=1e-4, eps=1e-8, num_replicas=N, peer_grads=peers)
        losses.append(loss + float(out["loss"].detach()) / N)
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--replicas", type=int, nargs="+", default=[2, 8])
    a = ap.parse_args()
    for N in a.replicas:
        exact, b16, pert = run(N, a.steps, "exact"), run(N, a.steps, "bf16"), run(N, a.steps, "perturbed")
        rel = lambda x, y: [abs(p - q) / abs(q) for p, q in zip(x, y)]
        print(f"N = {N}: reduced loss per step")
        print("  fp32 / exact buckets :", " ".join(f"{x:10.4f}" for x in exact))
        print("  bf16 buckets         :", " ".join(f"{x:10.4f}" for x in b16), "   rel. diff", " ".join(f"{x:.1e}" for x in rel(b16, exact)))
        print("  1e-5 weight perturb. :", " ".join(f"{x:10.4f}" for x in pert), "   rel. diff", " ".join(f"{x:.1e}" for x in rel(pert, exact)))


if __name__ == "__main__":
    main()
