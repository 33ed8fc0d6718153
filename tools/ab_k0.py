"""A/B of the mask-build variants (K0): round-1 global RED.AND vs the two-pass segment build
(K0b) at several group sizes, with and without row lengths, int64 and int32 ids.

Usage (GPU box): python tools/ab_k0.py [cfg] [batch] [variants,comma]
  variant = red | bucket[:group] | cluster[:cs[:groups]] ; default: red,bucket,bucket:8,bucket:16,bucket:32,bucket:64,bucket:100000
Prints ms of the mask phase (CUDA events inside the library), GB/s on the VALID id bytes
(sum T_b * 8 + B*W*4) and the fraction of the measured HBM peak; checks that every variant yields the
same nnz(M) and rewards as the first one.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    variants = (sys.argv[3] if len(sys.argv) > 3 else
                "red,bucket,bucket:8,bucket:16,bucket:32,bucket:64,bucket:100000").split(",")
    peak, _ = bench.measured_peak()
    t0 = time.time()
    pb = synth.make_problem(cfg, 1.0)
    coo = pb.a.tocoo()
    ctx = SpaiContext(pb.n, pb.edge_row, pb.edge_col, pb.edge_val, coo.row, coo.col, coo.data, device=0)
    dev = torch.device("cuda", 0)
    acts, lens = bench.device_trajectories(pb.num_edges, batch, 0, dev, 0.5)
    torch.cuda.synchronize()
    W = (pb.num_edges + 31) // 32
    valid = float(lens.sum()) * 8 + batch * W * 4.0
    padded = float(acts.numel()) * 8 + batch * W * 4.0
    print(f"{cfg}: n={pb.n} E={pb.num_edges} B={batch} T={acts.shape[1]} valid ids {int(lens.sum())} "
          f"({valid / 1e9:.2f} GB valid, {padded / 1e9:.2f} GB padded) built in {time.time() - t0:.1f}s", flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ctx.enable_timing(True)
    ref = None
    rows = []
    a32 = acts.to(torch.int32)
    for v in variants:
        name, _, grp = v.partition(":")
        os.environ["SPAI_K0_VARIANT"] = name
        for k in ("SPAI_K0B_GROUP", "SPAI_K0C_CS", "SPAI_K0C_GROUPS"):
            os.environ.pop(k, None)
        if name == "cluster" and grp:                      # cluster:<cs>:<groups>
            cs, _, nh = grp.partition(":")
            if cs and cs != "0":
                os.environ["SPAI_K0C_CS"] = cs
            if nh:
                os.environ["SPAI_K0C_GROUPS"] = nh
        elif grp:
            os.environ["SPAI_K0B_GROUP"] = grp
        for label, a_in, ln in (("int64+len", acts, lens), ("int64 padded", acts, None), ("int32+len", a32, lens)):
            if name == "red" and label == "int32+len":
                continue
            ms = []
            for it in range(4):
                flush.zero_()
                out = ctx.reward_batch(a_in, 0.5, "copy", torch.float32, lengths=ln)
                torch.cuda.synchronize()
                if it:
                    ms.append(ctx.last_timing().ms_masks)
            tm = ctx.last_timing()
            got = (out["nnz_m"].cpu().numpy(), out["reward"].cpu().numpy())
            if ref is None:
                ref = got
            same = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
            m = float(np.median(ms))
            byt = valid if ln is not None else padded
            if a_in.dtype == torch.int32:
                byt = float(lens.sum()) * 4 + batch * W * 4.0
            row = {"cfg": cfg, "B": batch, "variant": v, "input": label, "k0_ms": m, "GBps_on_read_bytes": byt / m / 1e6,
                   "frac_hbm_peak": byt / m / 1e6 / peak, "GBps_on_valid_int64_bytes": valid / m / 1e6,
                   "step_ms": tm.ms_total, "reward_ms": tm.ms_reward, "transpose_ms": tm.ms_transpose, "same_as_first": same}
            rows.append(row)
            print(json.dumps(row), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
