"""Pure-write, pure-read and copy bandwidth of this B200's HBM with plain library kernels (torch fill_ / sum / copy_), as a yardstick
for the stash streams: mlp_fwd_tc / mlp_dgrad_tc are write-only streams (4.0 / 3.8 GB per fine pass), mlp_wgrad_tc is read-only (8.2 GB)."""
import json, torch
dev = torch.device("cuda")
out = {}
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for gb in (1, 4, 8):
    n = gb * (1 << 30) // 4
    x = torch.empty(n, dtype=torch.float32, device=dev)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    t = timed(lambda: x.fill_(1.0)); out["write_%dGB" % gb] = round(gb * 1.073741824 / t, 1)
    t = timed(lambda: x.zero_()); out["memset_%dGB" % gb] = round(gb * 1.073741824 / t, 1)
    t = timed(lambda: x.sum()); out["read_%dGB" % gb] = round(gb * 1.073741824 / t, 1)
    t = timed(lambda: y.copy_(x)); out["copy_%dGB_rw" % gb] = round(2 * gb * 1.073741824 / t, 1)
    del x, y
print(json.dumps(out))
