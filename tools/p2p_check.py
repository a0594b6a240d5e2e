"""Peer-to-peer sanity on a multi-GPU box: NVLink status and measured d2d copy bandwidth."""
import subprocess, torch
print(subprocess.run(["nvidia-smi", "nvlink", "--status", "-i", "0"], capture_output=True, text=True).stdout[:1500])
n = torch.cuda.device_count()
print("devices", n, "p2p", [[torch.cuda.can_device_access_peer(i, j) for j in range(n) if j != i] for i in range(n)])
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.device(0):
    e0.record()
    for _ in range(10):
        b.copy_(a)
    e1.record()
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("peer copy GB/s:", 10 * 256 / 1024 / (e0.elapsed_time(e1) / 1e3))
