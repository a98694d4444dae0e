"""Host (enqueue) time per step of the sharded layer and of its end-to-end pipeline: the Python + launch cost a rank
pays per step, measured on one GPU at world size 1 with the device far behind (products shape: 28 ms of kernels per
step), so perf_counter over the un-synchronised loop is the host's own time.  At 8 GPUs a step is 4.2 ms of device
time: whatever the host needs beyond that bounds the step."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from benchmarks.layer import ShardedLayerBench, SingleGpuLayerBench  # noqa: E402


def enqueue_ms(fn, k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / k * 1e3


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29655")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    out = {}
    for name, make in (("single_gpu_layer", lambda: SingleGpuLayerBench(bench.WORKLOADS["products"], dev)),
                       ("sharded_layer_world1", lambda: ShardedLayerBench(bench.WORKLOADS["products"], 0, 1, dev, calibrate=False))):
        r = make()
        for _ in range(3):
            r.step()
        rec = {"step_enqueue_ms": round(enqueue_ms(r.step, 8), 3)}

        def e2e_like():
            for p in r.params:
                p.grad = None
            y = r._layer(r.x)
            y.backward(r.gout)
            flat = torch.cat([p.grad.reshape(-1) for p in r.params] + [y.detach()[:: max(1, y.shape[0] // 1024)].sum().reshape(1)])
            return flat
        rec["step_plus_result_pack_enqueue_ms"] = round(enqueue_ms(e2e_like, 8), 3)
        out[name] = rec
        del r
        torch.cuda.empty_cache()
    json.dump(out, sys.stdout, indent=1)
    print()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
