"""How often do the inverse-CDF bin decisions of sample_pdf_merge_kernel differ from the oracle's (torch CPU summation order),
and what does that do to the rendered colour?  Scenes: the random-init colour field with an ANALYTIC density -- a spherical
shell of radius 1 around the origin, peak 300, thickness `thick` (0.3 soft ... 0.01 a thin surface: many bins with ~zero pdf
mass, where the den<1e-5 snap and the searchsorted ties live)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("fast-learning-nerf_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import nerf_oracle as O
from flnerf_b200 import ops, synthetic

def rays(B, seed):
    g = torch.Generator().manual_seed(seed)
    H = W = 800
    K = synthetic.intrinsics(H, W, 1111.11)
    poses = synthetic.lego_like_poses(8)
    ro, rd = [], []
    for i in range(8):
        o, d = O.camera_rays(H, W, K, torch.as_tensor(poses[i][:3, :4]).float())
        sel = torch.randint(0, H * W, (B // 8,), generator=g)
        ro.append(o.reshape(-1, 3)[sel]); rd.append(d.reshape(-1, 3)[sel])
    return O.pack_rays(H, W, K, torch.cat(ro), torch.cat(rd), 2.0, 6.0, ndc=False)

def study(B=2048, thick=0.1, det=True, seed=3):
    r = rays(B, seed)
    pc, pf = O.init_params(41), O.init_params(42)
    o, d, view = r[:, 0:3], r[:, 3:6], r[:, 8:11]
    z = O.coarse_depths(r[:, 6:7], r[:, 7:8], 64, False, None)
    def sigma(pts):
        return 300.0 * torch.exp(-((pts.norm(dim=-1) - 1.0) / thick) ** 2)
    pts = o[:, None] + d[:, None] * z[..., None]
    raw = O.query(pc, pts, view)
    raw[..., 3] = sigma(pts)
    _, _, _, w, _ = O.composite(raw, z, d, None, True)
    u = None if det else torch.rand(B, 128, generator=torch.Generator().manual_seed(seed + 1))
    zo, zso = O.fine_depths(z, w, 128, u)
    zk, zsk, _ = ops.sample_pdf_merge(z.cuda(), w.cuda(), 128, det, None if det else u.cuda())
    zk, zsk = zk.cpu(), zsk.cpu()
    err = (zsk - zso).abs()
    def fine_rgb(zz):
        ptf = o[:, None] + d[:, None] * zz[..., None]
        rawf = O.query(pf, ptf, view)
        rawf[..., 3] = sigma(ptf)
        return O.composite(rawf, zz, d, None, True)[0]
    rgb_o, rgb_k = fine_rgb(zo), fine_rgb(zk)
    bad_rays = (err > 3e-5).any(-1)
    return {"B": B, "thick": thick, "det": det, "samples_off_3e-5": float((err > 3e-5).double().mean()), "max_dz": float(err.max()),
            "rays_touched": int(bad_rays.sum()), "acc_mean": float(w.sum(-1).mean()),
            "rgb_max_abs": float((rgb_o - rgb_k).abs().max()), "rgb_rel_l2": float((rgb_o - rgb_k).norm() / rgb_o.norm())}

if __name__ == "__main__":
    with torch.no_grad():
        for thick in (0.3, 0.05, 0.01):
            for det in (True, False):
                print(json.dumps(study(int(os.environ.get("PF_RAYS", 2048)), thick, det)), flush=True)
