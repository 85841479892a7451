"""NVLink peer bandwidth between GPU 0 and GPU 1 of one box (single process): copy-engine
copies each way, both ways at once, and an SM kernel reading peer memory (torch add).
    python tools/p2p_bw.py"""
import json
import torch

assert torch.cuda.device_count() >= 2
n = 1 << 28   # 1 GiB of fp32
a0 = torch.empty(n, device="cuda:0")
a1 = torch.empty(n, device="cuda:1")
b0 = torch.empty(n, device="cuda:0")
b1 = torch.empty(n, device="cuda:1")
print(json.dumps({"can_access_peer_0_1": torch.cuda.can_device_access_peer(0, 1)}))


def timed(fn, dev, reps=5):
    fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    with torch.cuda.device(dev):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    return s.elapsed_time(e) / reps


gb = n * 4 / 1e9
t = timed(lambda: a1.copy_(a0, non_blocking=True), 0)
print(json.dumps({"copy 0->1 GB/s": gb / t * 1e3}))
t = timed(lambda: a0.copy_(a1, non_blocking=True), 0)
print(json.dumps({"copy 1->0 GB/s": gb / t * 1e3}))
s1 = torch.cuda.Stream(device=1)


def both():
    a1.copy_(a0, non_blocking=True)
    with torch.cuda.stream(s1):
        b0.copy_(b1, non_blocking=True)


t = timed(both, 0)
torch.cuda.synchronize(1)
print(json.dumps({"bidirectional, per direction GB/s (lower bound)": gb / t * 1e3}))
