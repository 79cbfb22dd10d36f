"""Times k_allreduce_oneshot alone (pipelined mode 1, then a drain) under torchrun:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/probe_allreduce.py
Reports per-call device time with and without an L2 flush in between, next to NCCL's."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qml_essentials_b200 import backend  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm

    lib = backend.load_library()
    n = 646
    nbytes = int(lib.qmlb_allreduce_buffer_bytes(n))
    sbuf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    sbuf.zero_()
    hdl = symm.rendezvous(sbuf, dist.group.WORLD)
    torch.cuda.synchronize()
    dist.barrier()
    ptrs = (C.c_void_p * world)(*[int(q) for q in hdl.buffer_ptrs])
    x = torch.full((n,), float(rank + 1), dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    out = {}
    for label, do_flush in (("no_flush", False), ("flush", True)):
        for _ in range(5):
            lib.qmlb_allreduce_peer(ptrs, world, rank, n, x.data_ptr(), y.data_ptr(), 1, st)
        dist.barrier()
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            if do_flush:
                flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            lib.qmlb_allreduce_peer(ptrs, world, rank, n, x.data_ptr(), y.data_ptr(), 1, st)
            e.record()
            ts.append((s, e))
        lib.qmlb_allreduce_peer(ptrs, world, rank, n, x.data_ptr(), y.data_ptr(), 2, st)
        torch.cuda.synchronize()
        assert abs(float(y[0]) - world * (world + 1) / 2) < 1e-12, float(y[0])
        ms = sorted(s.elapsed_time(e) for s, e in ts)
        out[label] = {"median_us": 1e3 * ms[len(ms) // 2], "min_us": 1e3 * ms[0], "max_us": 1e3 * ms[-1]}
        dist.barrier()
    ts = []
    for _ in range(30):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dist.all_reduce(x)
        e.record()
        ts.append((s, e))
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ts)
    out["nccl_flush"] = {"median_us": 1e3 * ms[len(ms) // 2], "min_us": 1e3 * ms[0]}
    if rank == 0:
        print(json.dumps({"world": world, "fence": os.environ.get("QMLB_AR_FENCE", "light"), **out}))
    dist.barrier()
    dist.destroy_process_group()


main()
