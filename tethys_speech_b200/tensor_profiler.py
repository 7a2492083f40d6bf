"""Tiresias-style tensor-size profiler — the side-car of speech_jobs/whisper_dist_tensorsize.py (WT:20-458).

Same surface and the same log files as the reference's `TensorProfiler` (tensor_sizes.txt, memory_usage.txt, summary.txt,
tiresias_tensorsize.txt, final_summary.json, tiresias_result.json, legacy_skewness_result.txt): per step it adds up the bytes of
every tensor handed to `log_tensor_size`, the "Tiresias tensorsize" is the mean step size after a warm-up of min(3, steps // 4)
steps (WT:207-222), the skewness numbers are `scipy.stats.skew` of the tensor sizes (model-wide, per operation, per tensor type;
WT:224-289). Host-side bookkeeping only — it never touches the step's kernels. The reference logs from inside its Keras layers
(TensorLoggingMixin, WT:461-476); here `profile_step` walks what the native program exposes: the batch, every named activation
buffer of the program (ts_*_get_buffer), the gradient arena and the parameters, variable by variable."""
import json
import os

import numpy as np

try:
    from scipy import stats as _stats
except Exception:  # pragma: no cover - scipy is in the image; keep the profiler usable without it
    _stats = None


def _skew(values):
    v = np.asarray(values, dtype=np.float64)
    if _stats is not None:
        return float(_stats.skew(v))
    m = v.mean()
    s2 = ((v - m) ** 2).mean()
    return float(((v - m) ** 3).mean() / s2 ** 1.5) if s2 > 0 else 0.0


class TensorProfiler:
    def __init__(self, log_dir="/workspace/tensor_logs", model_name="whisper_small", verbose=True):
        self.log_dir, self.model_name, self.verbose = log_dir, model_name, verbose
        self.current_step = 0
        self.current_step_size = 0
        self.step_tensor_sizes = []
        self.operation_tensor_sizes = {}
        self.tensor_details = []
        self.memory_usage = []
        os.makedirs(log_dir, exist_ok=True)
        self.tensor_log_file = open(os.path.join(log_dir, "tensor_sizes.txt"), "w")
        self.tensor_log_file.write("step,operation,tensor_type,size_bytes,size_mb,shape\n")
        self.memory_log_file = open(os.path.join(log_dir, "memory_usage.txt"), "w")
        self.memory_log_file.write("step,gpu_memory_mb,cpu_memory_mb\n")
        self.summary_log_file = open(os.path.join(log_dir, "summary.txt"), "w")
        self.summary_log_file.write("step,total_tensor_size_mb,num_operations,avg_tensor_size_mb\n")
        self.tiresias_log_file = open(os.path.join(log_dir, "tiresias_tensorsize.txt"), "w")
        self.tiresias_log_file.write("step,tensorsize_mb\n")

    # -- WT:426-447 ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _calculate_tensor_size(tensor):
        if tensor is None:
            return 0
        if hasattr(tensor, "element_size") and hasattr(tensor, "numel"):          # torch
            return int(tensor.numel() * tensor.element_size())
        if hasattr(tensor, "nbytes"):                                             # numpy
            return int(tensor.nbytes)
        shape = getattr(tensor, "shape", None)
        n = int(np.prod([int(d) for d in shape])) if shape is not None else 1
        return n * int(getattr(getattr(tensor, "dtype", None), "size", 4) or 4)   # WT:443: 4 bytes when the dtype is unknown

    # -- WT:55-98 ------------------------------------------------------------------------------------------------------
    def log_tensor_size(self, tensor, name, tensor_type="activation"):
        if tensor is None:
            return 0
        size_bytes = self._calculate_tensor_size(tensor)
        size_mb = size_bytes / (1024 * 1024)
        shape = [int(d) for d in tensor.shape] if hasattr(tensor, "shape") else "unknown"
        self.current_step_size += size_bytes
        self.operation_tensor_sizes.setdefault(name, []).append(size_bytes)
        self.tensor_details.append({"step": self.current_step, "operation": name, "tensor_type": tensor_type, "size_bytes": size_bytes,
                                    "size_mb": size_mb, "shape": shape})
        self.tensor_log_file.write(f"{self.current_step},{name},{tensor_type},{size_bytes},{size_mb:.4f},{shape}\n")
        self.tensor_log_file.flush()
        return size_bytes

    def log_gradients(self, gradients, variables):                      # WT:100-105
        for i, (grad, var) in enumerate(zip(gradients, variables)):
            if grad is not None:
                self.log_tensor_size(grad, f"gradient_{getattr(var, 'name', None) or (var if isinstance(var, str) else f'variable_{i}')}", "gradient")

    def log_model_parameters(self, model):                              # WT:107-130
        total = 0
        names = getattr(model, "variable_names", None) or [f"variable_{i}" for i in range(len(model.trainable_variables))]
        for name, var in zip(names, model.trainable_variables):
            total += self.log_tensor_size(var, f"param_{name}", "parameter")
        return {"total_params_bytes": total, "trainable_params_bytes": total, "total_params_mb": total / (1024 * 1024)}

    def log_memory_usage(self):                                         # WT:132-178
        gpu = cpu = 0.0
        try:
            import torch

            if torch.cuda.is_available():
                gpu = torch.cuda.memory_allocated() / (1024 * 1024)
        except Exception:
            pass
        try:
            import psutil

            cpu = psutil.Process().memory_info().rss / (1024 * 1024)
        except Exception:
            pass
        info = {"step": self.current_step, "gpu_memory_mb": gpu, "cpu_memory_mb": cpu}
        self.memory_usage.append(info)
        self.memory_log_file.write(f"{self.current_step},{gpu:.2f},{cpu:.2f}\n")
        self.memory_log_file.flush()
        return info

    def start_step(self, step):                                         # WT:180-184
        self.current_step = step
        self.current_step_size = 0
        if self.verbose:
            print(f"📊 Step {step} 텐서 프로파일링 시작")

    def end_step(self):                                                 # WT:186-205
        step_mb = self.current_step_size / (1024 * 1024)
        self.step_tensor_sizes.append(step_mb)
        n_ops = sum(1 for d in self.tensor_details if d["step"] == self.current_step)
        avg = step_mb / n_ops if n_ops else 0
        self.summary_log_file.write(f"{self.current_step},{step_mb:.4f},{n_ops},{avg:.4f}\n")
        self.summary_log_file.flush()
        self.tiresias_log_file.write(f"{self.current_step},{step_mb:.4f}\n")
        self.tiresias_log_file.flush()
        if self.verbose:
            print(f"📊 Step {self.current_step} 완료 - TensorSize: {step_mb:.2f} MB")
        return step_mb

    def get_tiresias_tensorsize(self):                                  # WT:207-222
        if not self.step_tensor_sizes:
            return 0
        warm = min(3, len(self.step_tensor_sizes) // 4)
        stable = self.step_tensor_sizes[warm:]
        return float(np.mean(stable if stable else self.step_tensor_sizes))

    def calculate_tensor_skewness(self):                                # WT:224-244
        sizes = [d["size_mb"] for d in self.tensor_details if d["size_bytes"] > 0]
        return _skew(sizes) if len(sizes) >= 3 else 0.0

    def calculate_operation_skewness(self):                             # WT:246-261
        return {op: _skew([s / (1024 * 1024) for s in sizes]) for op, sizes in self.operation_tensor_sizes.items() if len(sizes) >= 3}

    def calculate_layer_type_skewness(self):                            # WT:263-289
        by_type = {}
        for d in self.tensor_details:
            if d["size_mb"] > 0:
                by_type.setdefault(d["tensor_type"], []).append(d["size_mb"])
        return {t: _skew(v) for t, v in by_type.items() if len(v) >= 3}

    def get_skewness_summary(self):                                     # WT:291-321
        sizes = [d["size_mb"] for d in self.tensor_details if d["size_mb"] > 0]
        return {"model_skewness": self.calculate_tensor_skewness(), "operation_skewness": self.calculate_operation_skewness(),
                "layer_type_skewness": self.calculate_layer_type_skewness(), "tensor_count": len(sizes),
                "mean_tensor_size_mb": float(np.mean(sizes)) if sizes else 0, "std_tensor_size_mb": float(np.std(sizes)) if sizes else 0,
                "min_tensor_size_mb": float(np.min(sizes)) if sizes else 0, "max_tensor_size_mb": float(np.max(sizes)) if sizes else 0}

    def get_summary(self):                                              # WT:360-394
        if not self.step_tensor_sizes:
            return {}
        sk = self.get_skewness_summary()
        s = self.step_tensor_sizes
        return {"total_steps": len(s), "tiresias_tensorsize_mb": self.get_tiresias_tensorsize(), "avg_step_tensorsize_mb": float(np.mean(s)),
                "max_step_tensorsize_mb": float(np.max(s)), "min_step_tensorsize_mb": float(np.min(s)), "std_step_tensorsize_mb": float(np.std(s)),
                "total_operations": len(self.tensor_details), "step_tensor_sizes": list(s), "model_skewness": sk["model_skewness"],
                "skewness_analysis": sk,
                "operation_stats": {op: {"total_size_mb": sum(v) / (1024 * 1024), "avg_size_mb": float(np.mean(v)) / (1024 * 1024), "count": len(v)}
                                    for op, v in self.operation_tensor_sizes.items()}}

    def save_final_results(self):                                       # WT:396-424
        summary = self.get_summary()
        with open(os.path.join(self.log_dir, "final_summary.json"), "w") as f:
            json.dump(summary, f, indent=2, default=str)
        with open(os.path.join(self.log_dir, "tiresias_result.json"), "w") as f:
            json.dump({"model": self.model_name, "tensorsize_mb": summary.get("tiresias_tensorsize_mb", 0), "skewness": summary.get("model_skewness", 0.0),
                       "total_steps": summary.get("total_steps", 0), "measurement_method": "Tiresias_style"}, f, indent=2)
        with open(os.path.join(self.log_dir, "legacy_skewness_result.txt"), "w") as f:
            f.write("model,skewness\n")
            f.write(f"{self.model_name},{summary.get('model_skewness', 0.0):.1f}\n")
        return summary

    def close(self):                                                    # WT:449-458
        for fh in (self.tensor_log_file, self.memory_log_file, self.summary_log_file, self.tiresias_log_file):
            try:
                fh.close()
            except Exception:
                pass


# what the native programs expose by name (ts_whisper_get_buffer / ts_w2v_get_buffer)
_BUFFERS = {"whisper": ("encoder_last_hidden_state", "last_hidden_state", "logits"),
            "w2v": ("extract_features", "hidden_states_in", "last_hidden_state", "quantized_features", "projected_states",
                    "projected_quantized_features", "contrastive_logits")}


def profile_step(profiler, step, model, inputs, step_fn):
    """One profiled train step (the body of the reference's profiled loop, WT: train_whisper_with_profiling): logs the batch, runs
    `step_fn()` (the product's train step, unchanged), then logs the program's named activation buffers, every variable's gradient
    and parameter, the memory in use, and closes the step. Returns (loss, step tensorsize in MB)."""
    profiler.start_step(step)
    for i, t in enumerate(inputs):
        if t is not None:
            profiler.log_tensor_size(t, f"input_{i}", "input")
    loss = step_fn()
    prog = model._prog
    for name in _BUFFERS["whisper" if prog.prefix == "ts_whisper" else "w2v"]:
        try:
            profiler.log_tensor_size(prog.buffer(name), name, "activation")
        except Exception:
            pass      # a buffer this head / mode does not have
    grads = [prog.view(prog.grads, n) for n in model.variable_names]
    profiler.log_gradients(grads, model.variable_names)
    profiler.log_model_parameters(model)
    profiler.log_memory_usage()
    return loss, profiler.end_step()
