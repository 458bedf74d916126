#!/usr/bin/env python
"""bench.py -- training rays/sec at 64 coarse + 128 fine samples (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

A "step" is one pass of the hot path over one batch: quadtree batch gather -> ray packing -> stratified depths ->
PE -> coarse MLP -> compositing -> inverse-CDF resampling + merge -> PE -> fine MLP -> compositing -> both MSE
losses (+ per-leaf max table) -> backward through both nets -> (gradient all-reduce) -> Adam.  Workload =
BASELINE config 2 ("lego 800x800, 64+128, N_rand=4096, quadtree on") on synthetic lego-like cameras/images;
under torchrun every rank takes 4096 rays of a N*4096-ray global batch (config 4), weak scaling.

`value`   : rays/s with rays, targets and the quadtree index buffer resident in HBM (CUDA events, max over ranks).
`e2e`     : the same metric through the reference-facing API (render() + loss.backward() + optimizer.step()) with
            HOST ray buffers: pinned H2D copy of (rays_o, rays_d, target) and a D2H read of the loss every step.
`roofline`: the dominant kernel (by measured time) against the roof that bounds it (measured bf16 tensor peak or measured
            HBM bandwidth; algorithmic FLOPs / bytes only), every MLP kernel's two fractions, and the whole step's tensor fraction.
`cpu_baseline` / --impl reference: the oracle port of the reference's PyTorch path on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))

FLOP_TRAIN_PER_RAY = 0.893190e9      # 256 MLP evaluations x 3 489 024 FLOP
FLOP_FWD_PER_SAMPLE = 1186816.0
FLOP_DGRAD_PER_SAMPLE = 2.0 * 557696
FLOP_WGRAD_PER_SAMPLE = 2.0 * 593408
# algorithmic HBM bytes per sample: activation / gradient stash images are 64 KB per 128-row tile and layer
BYTES_FWD_PER_SAMPLE = (16384 + 10 * 65536 - 32768 + 36864) / 128.0 + 16      # PE tile in; 9.5 act slots + masks, raw out
BYTES_DGRAD_PER_SAMPLE = (36864 + 10 * 65536 - 32768) / 128.0 + 16            # masks + draw in; 9.5 gradient slots out
BYTES_WGRAD_PER_SAMPLE = (672 + 608 + 32) * 1024 / 128.0                      # dY slots (dY5 twice) + activations + PE


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(tf_burst=1590.0, tf_sust=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms from before the warm-up until after the e2e loop; every
    row is stamped on arrival and ``window(t0, t1)`` reports the rows that fell inside the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7]
        note = "timed region"
        if not rows:   # a region shorter than the sampling period: fall back to every row taken under load
            rows, note = [r for (_, r) in self.rows if len(r) >= 7], "warm-up + timed + e2e (no sample fell inside the timed region)"
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm), "window": note}


# ------------------------------------------------------------------------------------------------- reference arm
def oracle_rays_per_s(n_rand, steps, warmup, threads=None):
    """The reference's own CPU path (oracle port: render + 2 MSE + backward + Adam), host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import nerf_oracle as O
    from flnerf_b200 import synthetic  # pose helpers only (no GPU work)
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    H = W = 800
    K = synthetic.intrinsics(H, W, 1111.111)
    pc, pf = O.init_params(0), O.init_params(1)
    opt = O.AdamState(list(pc.values()) + list(pf.values()))
    rs = np.random.RandomState(0)
    times = []
    for it in range(warmup + steps):
        pose = torch.as_tensor(synthetic.pose_spherical(float(rs.uniform(-180, 180)), -30.0, 4.0)[:3, :4])
        o, d = O.camera_rays(H, W, K, pose)
        sel = torch.from_numpy(rs.choice(H * W, n_rand, replace=False))
        rays = O.pack_rays(H, W, K, o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], 2.0, 6.0, ndc=False)
        tgt = torch.from_numpy(rs.uniform(0, 1, (n_rand, 3)).astype(np.float32))
        t0 = time.perf_counter()
        O.train_step(rays, tgt, pc, pf, opt, 64, 128, white_bkgd=True, t_rand=torch.rand(n_rand, 64),
                     u=torch.rand(n_rand, 128), det_fine=False)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return n_rand / t, t, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rand = 1024
    v, t, cores = oracle_rays_per_s(n_rand, args.steps, args.warmup)
    sample = "%d-ray batches (of the 4096-ray step) x %d steps, oracle port of nerf-ours render+loss+backward+Adam, fp32" % (n_rand, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "training rays/sec (64+128 samples)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * 4096 / n_rand,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lego 800x800, 64+128, N_rand=4096 (timed on a %d-ray sample per step)" % n_rand},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import model
    import render as R
    import run_nerf
    import run_nerf_helpers as Hh
    import tree
    from flnerf_b200 import lib, ops, synthetic
    from flnerf_b200.engine import FusedAdam, Trainer

    H = W = 800
    focal = 1111.111
    K = synthetic.intrinsics(H, W, focal)
    n_img = args.images
    n_rand = args.n_rand                      # rays per rank per step
    poses = synthetic.lego_like_poses(n_img)
    torch.manual_seed(0)
    images = synthetic.render_scene(H, W, K, poses, n_samples=48, device=dev)
    mgr = tree.QuadTreeManager(H, W, K, images, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2,
                               max_level=7, device=dev, seed=0)

    def make(seed):
        torch.manual_seed(seed)
        return model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                          precision=args.precision).to(dev)
    nc, nf = make(0), make(1)
    opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
    tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, world_size=world, rank=rank)
    n_rays = mgr.emit_epoch(down_scale=1)
    gb = n_rand * world                       # global batch
    total = args.warmup + args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs
    first = 0
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        tr.step_from_tree(mgr, first, gb); first += gb
    barrier()
    lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss = tr.step_from_tree(mgr, first, gb); first += gb
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = gb * args.steps / (ms * 1e-3)
    loss_host = loss.tolist()

    # ---------------- e2e: reference-facing API with host buffers
    q = run_nerf.NetworkQuery(Hh.get_embedder(10)[0], Hh.get_embedder(4)[0], 65536)
    hb = []
    for i in range(3):      # three pinned host batches, rotated
        o, d, tg, _ = mgr.batch(first + rank, n_rand, world); first += gb
        hb.append([x.cpu().pin_memory() for x in (o, d, tg)])
    bytes_in = sum(x.numel() * 4 for x in hb[0])

    def api_step(i):
        ho, hd, ht = hb[i % 3]
        ro, rd, tg = ho.to(dev, non_blocking=True), hd.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
        rgb, _, _, ex = R.render(H, W, K, chunk=32768, rays=torch.stack([ro, rd], 0), ndc=False, near=2.0, far=6.0,
                                 use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64,
                                 N_importance=128, white_bkgd=True, perturb=1.0, raw_noise_std=0.0, retraw=True)
        opt.zero_grad()
        l_f, l_c = Hh.img2mse(rgb, tg), Hh.img2mse(ex["rgb0"], tg)
        (l_f + l_c).backward()
        if world > 1:
            tr.bucket.div_(world)
            dist.all_reduce(tr.bucket)
        opt.step()
        return torch.stack([l_f.detach(), l_c.detach()]).cpu()       # D2H read of the step's result (8 bytes)

    for i in range(max(1, min(args.warmup, 3))):
        api_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        api_step(i)
    f1.record()
    barrier()
    t = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = gb * args.steps / (float(t.item()) * 1e-3)
    clocks.stop()
    clk = clocks.window(w0, w1)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- eval path (informational): one full 800x800 frame through render() without gradients
    eval_info = None
    if args.precision == "bf16":
        with torch.no_grad():
            pose = mgr._poses_dev[0]
            kw = dict(chunk=32768, c2w=pose, ndc=False, near=2.0, far=6.0, use_viewdirs=True, network_query_fn=q, network_fn=nc,
                      network_fine=nf, N_samples=64, N_importance=128, white_bkgd=True, perturb=0.0, raw_noise_std=0.0)
            R.render(H, W, K, **kw)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            R.render(H, W, K, **kw)
            g1.record()
            torch.cuda.synchronize()
            ems = g0.elapsed_time(g1)
        eval_info = {"rays_per_s": H * W / (ems * 1e-3), "ms_per_frame": ems, "frame": "%dx%d, 64+128 samples, render() under no_grad" % (H, W),
                     "tflops": H * W * 256 * FLOP_FWD_PER_SAMPLE / (ems * 1e-3) / 1e12}
    # ---------------- per-kernel roofline (rank 0, fine pass: 4096 x 192 rows), CUDA events on the launch stream
    peaks = measured_peaks()
    roof, kernels = None, {}
    if args.precision == "bf16":
        o, d, tg, _ = mgr.batch(0, n_rand, 1)
        r11 = ops.pack_rays(o, d, 2.0, 6.0, False, H, W, focal)
        z = ops.coarse_depths(r11, 192, True, False, None, 3, 0)
        n = n_rand * 192
        tiles, dirpe = ops.encode_tc(r11, z)
        flat, packed = nf._weights()
        raw, stash = ops.mlp_forward(ops.MODE_BF16, flat, packed, tiles, dirpe, n, 192, True)
        draw = torch.randn(n, 4, device=dev) * 1e-3
        gbuf = torch.zeros_like(flat)
        ws = ops.mlp_backward(ops.MODE_BF16, flat, packed, tiles, dirpe, stash, draw, gbuf, n, 192)
        stash_l = ops._alloc_bytes(lib.load().flnerf_mlp_stash_bytes(1, n, 192, 1), dev)
        import ctypes as C

        def fwd():
            lib.check(lib.load().flnerf_mlp_forward(ops._ctx(raw), 1, ops._ptr(flat), ops._ptr(packed), n, 192, ops._ptr(tiles),
                                                    ops._ptr(dirpe), ops._ptr(raw), ops._ptr(stash_l), 1, ops._stream()), "fwd")
        cases = {"mlp_fwd_tc": (fwd, FLOP_FWD_PER_SAMPLE, BYTES_FWD_PER_SAMPLE),
                 "mlp_dgrad_tc": (lambda: ops.mlp_backward(1, flat, packed, tiles, dirpe, stash, draw, gbuf, n, 192, 1, ws),
                                  FLOP_DGRAD_PER_SAMPLE, BYTES_DGRAD_PER_SAMPLE),
                 "mlp_wgrad_tc": (lambda: ops.mlp_backward(1, flat, packed, tiles, dirpe, stash, draw, gbuf, n, 192, 2, ws),
                                  FLOP_WGRAD_PER_SAMPLE, BYTES_WGRAD_PER_SAMPLE)}
        for name, (fn, flop, nbytes) in cases.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b) / 5 * 1e-3
            tf, gbs = n * flop / dt / 1e12, n * nbytes / dt / 1e9
            kernels[name] = {"ms": dt * 1e3, "tflops": tf, "tensor_frac": tf / peaks["tf_sust"], "gbs": gbs,
                             "hbm_frac": gbs / peaks["hbm"]}
        top = max(kernels, key=lambda k: kernels[k]["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get(top)
        # the dominant kernel is judged against the roof that bounds it: whichever of its two fractions is larger
        k = kernels[top]
        if k["hbm_frac"] > k["tensor_frac"]:
            roof = {"bound": "hbm", "kernel": top, "achieved": k["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": k["hbm_frac"], "traffic": traffic,
                    "peak_source": peaks["src"] + " HBM copy bandwidth (a read-mostly stream can sit slightly above it)"}
        else:
            roof = {"bound": "tensor", "kernel": top, "achieved": k["tflops"], "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                    "frac": k["tensor_frac"], "traffic": traffic,
                    "peak_source": peaks["src"] + " bf16 sustained (kernel timed inside a long step)"}
        roof["kernels"] = kernels
        roof["step"] = {"bound": "tensor", "achieved": value / world * FLOP_TRAIN_PER_RAY / 1e12, "unit": "TFLOP/s",
                        "peak": peaks["tf_sust"], "frac": value / world * FLOP_TRAIN_PER_RAY / 1e12 / peaks["tf_sust"]}
    # ---------------- CPU baseline (oracle port) on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, tcpu, cores = oracle_rays_per_s(512, 3, 1)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": "512-ray batches x 3 steps (1 warm-up) of the same 64+128 step, oracle port on host cores"}
    print(json.dumps({
        "metric": "training rays/sec (64+128 samples)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "lego 800x800 (synthetic cameras/images, %d train views), 64+128 samples, N_rand=%d per GPU "
                               "(global %d), quadtree on (init_level 2)" % (n_img, n_rand, gb),
                   "parallelism": "ray-sharded data parallel x%d, one gradient all-reduce per step" % world,
                   "l2": "per-step working set (activation stash ~%.1f GB) exceeds the 126 MB L2" % (n_rand * 256 * 5.1e3 / 1e9),
                   "epoch_rays": n_rays},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 8},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        "eval_render": eval_info, "loss": loss_host}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("FLNERF_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--n_rand", type=int, default=4096)
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
