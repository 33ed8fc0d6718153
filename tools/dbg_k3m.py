import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pb = synth.make_problem(cfg, 1.0)
coo = pb.a.tocoo()
dev = torch.device("cuda", 0)
acts, lens = bench.device_trajectories(pb.num_edges, batch, 0, dev, 0.5)
os.environ["SPAI_K3M_SPLIT"] = sys.argv[3] if len(sys.argv) > 3 else "3"
ctx = SpaiContext(pb.n, pb.edge_row, pb.edge_col, pb.edge_val, coo.row, coo.col, coo.data, device=0)
ctx.enable_timing(True)
for ntm in ("4", "2"):
    os.environ["SPAI_K3M_NTM"] = ntm
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["SPAI_K3M_DEBUG"] = str(dbg)
        ms = []
        for it in range(3):
            ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)
            torch.cuda.synchronize()
            if it: ms.append(ctx.last_timing().ms_reward)
        print(json.dumps({"ntm": ntm, "dbg(1=noMMA 2=noSTTM 4=noLDTM)": dbg, "reward_ms": float(np.median(ms))}), flush=True)
ctx.close()
