"""Raw PCIe throughput of this box (pinned host memory, one GPU): H2D alone, D2H alone, both at once; whole buffers and 4 MB chunks."""
import json, torch
n = 48 * 1024 * 1024      # doubles: 384 MB, the state of the 16M-cell mesh
h1 = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
gb = n * 8 / 1e9
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
def chunked(k=96):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d1[i*c:(i+1)*c].copy_(h1[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h2[i*c:(i+1)*c].copy_(d2[i*c:(i+1)*c], non_blocking=True)
out = {"h2d_GBs": gb / timed(h2d) * 1e3, "d2h_GBs": gb / timed(d2h) * 1e3, "both_each_GBs": gb / timed(both) * 1e3, "both_chunked_4MB_each_GBs": gb / timed(chunked) * 1e3}
print(json.dumps(out))
