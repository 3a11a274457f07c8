"""Pure-write HBM ceiling: fill / memset of buffers larger than L2 (CUDA events)."""
import json
import torch
dev = torch.device("cuda:0")
for mb in (194, 512, 2048):
    n = mb * 1000 * 1000 // 4
    bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(4)]
    for fn_name in ("fill_", "zero_"):
        for b in bufs:
            getattr(b, fn_name)(*([1.0] if fn_name == "fill_" else []))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            for b in bufs:
                getattr(b, fn_name)(*([1.0] if fn_name == "fill_" else []))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(json.dumps({"op": fn_name, "MB": mb, "us": round(us, 1), "write_GBps": round(n * 4 / us / 1e3, 1)}))
