"""Runs bench.main() with torch.cuda and rome_b200.Context replaced by inert stand-ins (see test_bench_dryrun.py).
Not a test module itself: started as a subprocess, once per rank, with RANK / WORLD_SIZE / MASTER_* in the environment
for the multi-rank dry run (process group on gloo, tensors on the CPU)."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 9.0  # ms for the K timed steps


class _Graph:
    def replay(self):
        pass


class _Ctx:
    """inert stand-in for rome_b200.Context: accepts every call bench.py makes, fills host outputs with ones"""
    def __init__(self, device=0):
        self.device = device
        self.launch_count = 0

    def __getattr__(self, name):
        def call(*a, **k):
            if name.startswith("eval_host"):
                for key in ("res", "stats"):
                    if isinstance(k.get(key), torch.Tensor):
                        k[key].fill_(1.0)
            if name in ("eval", "peer_signal", "peer_wait", "peer_barrier", "push_halo"):
                self.launch_count += 1
            if name == "ipc_export":
                return b"\0" * 64
            if name == "particles_device":
                return (1 << 21, 1296, 48, 0, 100, 104)
            if name == "memcpy_d2h":
                a[0][...] = 0
                return None
            if name == "peer_gave_up":
                return False
            return 1 << 21 if name in ("malloc_device", "peer_state_alloc", "ipc_import") else None
        return call


def install():
    import torch.distributed as dist
    import bench
    import rome_b200 as rb

    def on_cpu(fn):
        def wrapped(*a, **k):
            k.pop("device", None)
            return fn(*a, **k)
        return wrapped

    for name in ("zeros", "randn", "tensor", "as_tensor"):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    torch.Tensor.pin_memory = lambda self: self
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.Stream, torch.cuda.Event, torch.cuda.CUDAGraph = _Stream, _Event, _Graph
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.graph = lambda g, stream=None: contextlib.nullcontext()
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **k: real_init("gloo")
    rb.Context = _Ctx

    return bench


if __name__ == "__main__":
    install().main()
