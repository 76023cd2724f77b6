"""Runs bench.py's `ours` arm WITHOUT a GPU: torch.cuda is stubbed, "cuda" devices become CPU tensors and
semtools_b200.capi is replaced by tests/fake_capi.py (oracle-backed).  What it checks is bench.py itself --
argument plumbing, section order, collectives (gloo, via STB_BENCH_ONE_GPU=1), deadlines, the shape of the
JSON lines.  Every number it prints is meaningless.  Launched by tests/test_bench_sim.py:

    python tests/bench_sim.py --gpus 1 --rows 30000 ...
    STB_BENCH_ONE_GPU=1 python -m torch.distributed.run --nproc-per-node 2 ... tests/bench_sim.py --gpus 2 ...
"""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

_real_torch = torch


class _Stream:
    cuda_stream = 0x5717

    def __init__(self, *a, **k):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


_cuda = types.SimpleNamespace(Stream=_Stream, Event=_Event, set_stream=lambda s: None, set_device=lambda d: None,
                              synchronize=lambda d=None: None, empty_cache=lambda: None, is_available=lambda: False)


def _device(*a, **k):
    if a and (a[0] == "cuda" or (isinstance(a[0], str) and a[0].startswith("cuda"))):
        return _real_torch.device("cpu")
    return _real_torch.device(*a, **k)


class _TorchProxy(types.ModuleType):
    """The real torch with `cuda` and `device` swapped; everything else passes through."""

    def __getattr__(self, name):
        if name == "cuda":
            return _cuda
        if name == "device":
            return _device
        return getattr(_real_torch, name)


_real_torch.Tensor.pin_memory = lambda self, *a, **k: self          # no accelerator here
proxy = _TorchProxy("torch")
proxy.__dict__["__path__"] = _real_torch.__path__
proxy.__dict__["__spec__"] = _real_torch.__spec__
sys.modules["torch"] = proxy

import semtools_b200  # noqa: E402
import fake_capi  # noqa: E402

semtools_b200.capi = fake_capi
sys.modules["semtools_b200.capi"] = fake_capi

import bench  # noqa: E402

if __name__ == "__main__":
    bench.main()
