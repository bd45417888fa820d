"""Test helper (not a product path): runs bench.py's single-rank measurement code on the CPU kernel emulator by
stubbing the handful of torch.cuda calls it makes, so that the assembly of the JSON line (metric, roofline, e2e,
clocks, cpu_baseline, decomposition, ...) is exercised in the GPU-less build container.  The numbers are
meaningless; tests/test_bench_model.py only checks structure.  Usage: python bench_on_emulator.py <emulator .so> <n>"""
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def main():
    emu, n = sys.argv[1], int(sys.argv[2])
    os.environ.setdefault("LAPS_BENCH_PARITY_N", "16")   # the gate's logic, not its size, is what this helper exercises
    from laps_b200 import capi
    capi.DEFAULT_LIB = emu
    _load = capi.load
    capi.load = lambda path=None: _load(path or emu)
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.ExternalStream = lambda ptr, device=None: None
    torch.cuda.Event = _Event
    torch.Tensor.pin_memory = lambda self: self
    import contextlib
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    _empty = torch.empty
    torch.empty = lambda *a, **k: _empty(*a, **{kk: v for kk, v in k.items() if kk != 'device'})
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # multi-rank: gloo in place of NCCL, host tensors in place of device ones
        import torch.distributed as dist
        _init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **kw: _init("gloo")
        torch.Tensor.cuda = lambda self, *a, **k: self
        _tensor = torch.tensor
        torch.tensor = lambda *a, **k: _tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sys.argv = ["bench.py", "--n", str(n), "--steps", "2", "--warmup", "3", "--cpu-n", "16", "--gpus", os.environ.get("WORLD_SIZE", "1")]
    return bench.main()


if __name__ == "__main__":
    sys.exit(main())
