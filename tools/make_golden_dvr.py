"""Generate tests/golden/dvr_*.npz from the REFERENCE's own CUDA kernels.

Run on a GPU box (the reference extensions are compiled from /root/reference by
oracle/build_ref.py in the build container and travel as oracle/_ref/*.so):

    gpurun -- python tools/make_golden_dvr.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each file holds the inputs and what third_lib/dvr, third_lib/dvxlr, third_lib/dvxlr_v2
returned for them.  tests/test_golden_dvr.py pins oracle/dvr_ref.c against these.
`dvr.render`'s grad_sigma is stored too but is a racy sum in the reference (dvr.cu:621-622).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests.inputs import dvr_inputs_cfg1, dvr_inputs_lidar, dvr_inputs_outside, dvr_inputs_ties  # noqa: E402

CASES = {
    "cfg1_int": lambda: dvr_inputs_cfg1(seed=0, integer_origin=True),
    "cfg1_frac": lambda: dvr_inputs_cfg1(seed=0, integer_origin=False),
    "lidar_small": lambda: dvr_inputs_lidar(M=1500, T=2, grid=(8, 64, 64), seed=5, pad=12),
    "outside": lambda: dvr_inputs_outside(zero_length=False),
    "ties": lambda: dvr_inputs_ties(M=6000),
}


def trim(lists, count):
    """keep only the used prefix of the [N,M,1026,...] lists"""
    k = int(count.max())
    return [x[:, :, :k] for x in lists], k


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dvr = build_ref.load("ref_dvr")
    dvxlr = build_ref.load("ref_dvxlr")
    dvxlr_v2 = build_ref.load("ref_dvxlr_v2")
    dev = torch.device("cuda:0")
    only = set(sys.argv[1:])
    for name, make in CASES.items():
        if only and name not in only:
            continue
        sigma, origin, points, tindex = make()
        rng = np.random.default_rng(123)
        regul = rng.standard_normal(sigma.shape).astype(np.float32)
        s, o, p, t, r = (torch.from_numpy(x).to(dev) for x in (sigma, origin, points, tindex, regul))
        grid = list(sigma.shape[1:])
        rec = dict(sigma=sigma, origin=origin, points=points, tindex=tindex, sigma_regul=regul)
        rec["occupancy"] = dvr.init(p, t, grid).cpu().numpy()
        for ph in ("test", "train"):
            a, b = dvr.render_forward(s, o, p, t, grid, ph)
            rec[f"fwd_{ph}_pred"], rec[f"fwd_{ph}_gt"] = a.cpu().numpy(), b.cpu().numpy()
        for loss in ("l1", "l2", "absrel"):
            a, b, g = dvr.render(s, o, p, t, loss)
            rec[f"render_{loss}_pred"], rec[f"render_{loss}_gt"] = a.cpu().numpy(), b.cpu().numpy()
            rec[f"render_{loss}_grad_racy"] = g.cpu().numpy()
        pred, gt, dd, idx, ray_pred, ind = dvxlr_v2.render_v2(s, o, p, t, r)
        p1, g1, dd1, idx1 = dvxlr.render(s, o, p, t)
        assert torch.equal(pred, p1) and torch.equal(dd, dd1) and torch.equal(idx, idx1)
        count = (ind >= 0).sum(-1).cpu().numpy()
        (dd_t, idx_t, rp_t, ind_t), k = trim([dd.cpu().numpy(), idx.cpu().numpy(), ray_pred.cpu().numpy(),
                                              ind.cpu().numpy()], count)
        rec.update(dvxlr_pred=pred.cpu().numpy(), dvxlr_gt=gt.cpu().numpy(), dvxlr_dd=dd_t,
                   dvxlr_idx=idx_t.astype(np.int16), dvxlr_ray_pred=rp_t, dvxlr_indicator=ind_t.astype(np.int8),
                   dvxlr_count=count.astype(np.int16), dvxlr_k=np.int32(k))
        gp = rng.standard_normal(pred.shape).astype(np.float32)
        grp = rng.standard_normal(tuple(ray_pred.shape)).astype(np.float32)
        em = torch.from_numpy(gp).to(dev)[..., None] * dd
        g_a, g_b = dvxlr_v2.get_grad_sigma_v2(em, idx, t, s, ind, torch.from_numpy(grp).to(dev))
        (g_c,) = dvxlr.get_grad_sigma(em, idx, t, s)
        rec.update(grad_pred=gp, grad_ray_pred_seed=np.int32(123), scatter_grad_sigma=g_a.cpu().numpy(),
                   scatter_grad_regul=g_b.cpu().numpy(), scatter_grad_sigma_v1=g_c.cpu().numpy())
        np.savez_compressed(os.path.join(out_dir, f"dvr_{name}.npz"), **rec)
        print(name, "ok; max list length", k, "rays hit", int((pred.cpu().numpy() >= 0).sum()))


if __name__ == "__main__":
    main()
