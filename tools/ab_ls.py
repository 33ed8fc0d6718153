"""Time reward modes (copy, ls = Householder, ls_gram = semi-normal equations) on one
problem in one process and report how far apart their rewards are.

Usage (GPU box): python tools/ab_ls.py [cfg] [batch] [modes,comma] [dtypes,comma]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    modes = (sys.argv[3] if len(sys.argv) > 3 else "ls,ls_gram").split(",")
    dtypes = (sys.argv[4] if len(sys.argv) > 4 else "f64,f32").split(",")
    t0 = time.time()
    pb = synth.make_problem(cfg, 1.0)
    coo = pb.a.tocoo()
    ctx = SpaiContext(pb.n, pb.edge_row, pb.edge_col, pb.edge_val, coo.row, coo.col, coo.data, device=0)
    acts = torch.from_numpy(synth.make_trajectories(pb.num_edges, batch, seed0=1000)).cuda()
    print(f"{cfg}: n={pb.n} E={pb.num_edges} B={batch} T={acts.shape[1]} built in {time.time() - t0:.1f}s", flush=True)
    ctx.enable_timing(True)
    ref = {}
    for dt in dtypes:
        tdt = torch.float64 if dt == "f64" else torch.float32
        for md in modes:
            out = ctx.reward_batch(acts, 0.5, md, tdt)
            torch.cuda.synchronize()
            ms, tot = [], []
            for _ in range(3):
                out = ctx.reward_batch(acts, 0.5, md, tdt)
                torch.cuda.synchronize()
                ms.append(ctx.last_timing().ms_reward)
                tot.append(ctx.last_timing().ms_total)
            rw = out["reward"].cpu().numpy()
            base = ref.setdefault("f64" if "f64" in dtypes else dt, rw)
            err = float(np.max(np.abs(rw - base) / np.maximum(1e-300, np.abs(base))))
            print(f"{md}/{dt}: whole step {np.median(tot):.3f} ms = {batch / np.median(tot) * 1e3:.3e} patterns/s; reward kernels {np.median(ms):.3f} ms -> {batch * pb.n / np.median(ms) * 1e3:.3e} row solves/s; "
                  f"max rel diff vs first f64 run {err:.2e}; nan={int(np.isnan(rw).sum())}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
