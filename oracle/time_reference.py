"""Time the REFERENCE ITSELF (PreconditionerEnv.update through oracle/ref_shim.py)
in the build container — it cannot travel to the GPU box. Writes
profiles/reference_cpu_container.json. TEST INFRASTRUCTURE.

    python oracle/time_reference.py
"""
from __future__ import annotations

import gc
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gflownet_spai_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle import spai_oracle as orc  # noqa: E402


def run(cfg, scale, batch, max_frac=0.5):
    import torch
    p = synth.make_problem(cfg, scale)
    coo = p.a.tocoo()
    acts = synth.make_trajectories(p.num_edges, batch, max_frac=max_frac)
    out = {"config": cfg, "scale": scale, "n": p.n, "num_edges": p.num_edges, "batch": batch,
           "threads": torch.get_num_threads(), "cpu_count": os.cpu_count()}
    t0 = time.perf_counter()
    ref = ref_shim.reference_update(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, acts, 0.5)
    out["reference_as_is_patterns_per_s"] = batch / (time.perf_counter() - t0)
    real_collect = gc.collect
    gc.collect = lambda *a, **k: 0                     # "reference-minus-GC" 
    for m in ref_shim.load_reference().modules:
        if hasattr(m, "gc"):
            m.gc.collect = gc.collect
    try:
        t0 = time.perf_counter()
        ref_shim.reference_update(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, acts, 0.5)
        out["reference_minus_gc_patterns_per_s"] = batch / (time.perf_counter() - t0)
    finally:
        gc.collect = real_collect
        for m in ref_shim.load_reference().modules:
            if hasattr(m, "gc"):
                m.gc.collect = real_collect
    t0 = time.perf_counter()
    port = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32),
                                 p.a.astype(np.float32), acts, 0.5, dtype=np.float32)
    out["oracle_port_patterns_per_s"] = batch / (time.perf_counter() - t0)
    out["max_abs_reward_diff_port_vs_reference"] = float(np.max(np.abs(port["reward"] - ref["reward"])))
    return out


def main():
    rows = [run("cfg1", 1.0, 32), run("cfg2", 0.25, 8), run("cfg2", 1.0, 2)]
    path = os.path.join(ROOT, "profiles", "reference_cpu_container.json")
    with open(path, "w") as f:
        json.dump({"note": "measured in the build container (no GPU), torch CPU; the reference reads an "
                           "undefined self.alpha (preconditioner.py:163): env.alpha is injected", "rows": rows}, f, indent=1)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
