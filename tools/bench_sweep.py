"""Config 5 of BASELINE.json: extract_fields N^3 SDF sweep (field.py:1286-1307) on the fused inference kernel.
Prints device time (CUDA events), end-to-end time including the single D2H copy, and the tensor-pipe fraction."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402
from nu_nerf_b200.sweep import extract_fields  # noqa: E402

M_SDF_HEAD = 459008


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).cuda()
    bmin, bmax = -torch.ones(3), torch.ones(3)
    extract_fields(bmin, bmax, 64, net.sdf_network.sdf)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    u = extract_fields(bmin, bmax, res, net.sdf_network.sdf, return_device=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    uh = extract_fields(bmin, bmax, res, net.sdf_network.sdf)
    wall = time.perf_counter() - t0
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1400.0
    tf = 2.0 * M_SDF_HEAD * res ** 3 / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": f"extract_fields {res}^3 SDF sweep (bf16 fused chain)", "device_ms": ms,
                      "end_to_end_s_with_d2h": wall, "points": res ** 3, "Mpts_per_s": res ** 3 / ms / 1e3,
                      "tflops": tf, "frac_of_tensor_peak": tf / peak, "inside_fraction": float((uh < 1.0).mean())}))


if __name__ == "__main__":
    main()
