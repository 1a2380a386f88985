#!/usr/bin/env python
"""Dry run of bench.py WITHOUT a GPU (test infrastructure): the CUDA library is replaced by its CPU emulation
(oracle/_build/librsq_emu.so, tests/cudaemu), torch.cuda by host stand-ins, NCCL by gloo — the library's rsq_comm_init / rsq_exchange by
the transport-free protocol over gloo.  It exercises every line of bench.py's control flow (pre-flight, the broadcast warm-up count,
timed steps, the strong-scaling leg, the e2e leg and its equality check, the per-rank gather) at toy sizes; the numbers mean nothing.

    python tools/bench_dryrun.py --workload chr1_10m --records 12000 --steps 2 --e2e-steps 1 --e2e-chunk 5000 --port-sample 8000 --cli-records 0
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/bench_dryrun.py --gpus 2 \
        --workload chr1_10m --records 8000 --steps 2 --min-warmup-s 0 --e2e-steps 1 --e2e-chunk 5000 --bed
"""
import os, sys, time, types, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RSQ_B200_LIB"] = os.path.join(ROOT, "oracle", "_build", "librsq_emu.so")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
class _Ev:
    def __init__(self, **k): self.t = 0
    def record(self, s=None): self.t = time.perf_counter()
    def elapsed_time(self, o): return (o.t - self.t) * 1e3
class _St:
    cuda_stream = 0
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.Stream = lambda: _St()
torch.cuda.set_stream = lambda s: None
torch.cuda.Event = _Ev
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **{kk: vv for kk, vv in k.items() if kk != "pin_memory"})
from rnaseqc_b200 import api
def _reduce(ptr, n):
    if n:
        buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n,))
        dist.all_reduce(torch.from_numpy(buf))
api.Backend.comm_unique_id = lambda self: b"\0" * 128
api.Context.comm_init = lambda self, cid, rank, world: self.set_partition(rank, world)
def _exchange(self):
    _reduce(*self.export_sample_counts()); _reduce(*self.export_counts()); self.import_counts()
api.Context.exchange = _exchange
_cs = api.Context.coverage_statistics
def _cov(self, ann, **kw):
    out = _cs(self, ann, **kw)
    if kw.get("owned") is not None:
        for k in out:
            t = torch.from_numpy(np.ascontiguousarray(out[k]).view(np.int64)); dist.all_reduce(t)
            out[k][:] = t.numpy().view(out[k].dtype)
    return out
api.Context.coverage_statistics = _cov
src = open(os.path.join(ROOT, "bench.py")).read()
src = src.replace('device = f"cuda:{local_rank}"', 'device = "cpu"').replace('device = f"cuda:{device_index}"', 'device = "cpu"').replace("n_genes=5234", "n_genes=150")
src = src.replace('dist.init_process_group("nccl", device_id=torch.device(device), timeout=datetime.timedelta(seconds=300))', 'dist.init_process_group("gloo")')
assert 'init_process_group("gloo")' in src and 'device = "cpu"' in src, "bench.py changed: update the patches of this dry run"
src = src.replace("n_records=2_000_000, quota=20000", "n_records=8000, quota=40")
mod = types.ModuleType("bench_dry"); mod.__file__ = os.path.join(ROOT, "bench.py")
exec(compile(src, os.path.join(ROOT, "bench.py"), "exec"), mod.__dict__)
mod.main()
