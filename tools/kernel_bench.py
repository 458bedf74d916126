"""Times the three tensor-core MLP kernels alone (fine pass: 4096 rays x 192 samples) with CUDA events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
import torch
import model
from flnerf_b200 import ops, lib

B, S = int(os.environ.get("KB_RAYS", 4096)), 192
MODE = int(os.environ.get("KB_MODE", 1))          # 1 = bf16, 2 = bf16x3
dev = torch.device("cuda")
torch.manual_seed(0)
net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16" if MODE == 1 else "bf16x3").to(dev)
rays = torch.cat([torch.randn(B, 3) * 0.3, torch.nn.functional.normalize(torch.randn(B, 3), dim=-1), 2 * torch.ones(B, 1),
                  6 * torch.ones(B, 1), torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)], -1).to(dev)
z = ops.coarse_depths(rays, S, True, False, None, 3, 0)
n = B * S
tiles, dirpe = ops.encode_tc(rays, z, MODE)
flat, packed = net._weights()
raw, stash = ops.mlp_forward(MODE, flat, packed, tiles, dirpe, n, S, True)
draw = torch.randn(n, 4, device=dev) * 1e-3
g = torch.zeros_like(flat)
ws = ops.mlp_backward(MODE, flat, packed, tiles, dirpe, stash, draw, g, n, S)
stash2 = ops._alloc_bytes(lib.load().flnerf_mlp_stash_bytes(MODE, n, S, 1), dev)
def fwd(train=1):
    lib.check(lib.load().flnerf_mlp_forward(ops._ctx(raw), MODE, ops._ptr(flat), ops._ptr(packed), n, S, ops._ptr(tiles),
                                            ops._ptr(dirpe), ops._ptr(raw), ops._ptr(stash2), train, ops._stream()), "fwd")
cases = {"fwd": (lambda: fwd(1), 1186816.0), "fwd_infer": (lambda: fwd(0), 1186816.0),
         "dgrad": (lambda: ops.mlp_backward(MODE, flat, packed, tiles, dirpe, stash, draw, g, n, S, 1, ws), 2.0 * 557696),
         "wgrad": (lambda: ops.mlp_backward(MODE, flat, packed, tiles, dirpe, stash, draw, g, n, S, 2, ws), 2.0 * 593408),
         "heads": (lambda: ops.mlp_backward(MODE, flat, packed, tiles, dirpe, stash, draw, g, n, S, 4, ws), 0.0)}
out = {}
for k, (fn, flop) in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    out[k] = (round(ms, 3), round(n * flop / ms / 1e9, 1))
print(os.environ.get("KB_TAG", ""), json.dumps(out))
