"""Micro-benchmark of nunerf_linear / nunerf_linear_dw alone (CUDA events, L2-exceeding operands)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import ops  # noqa: E402

DEV = "cuda"


def mk(M, w, planes):
    p = ops.P(M, w, planes, DEV)
    p.t.normal_()
    return p


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 384000
    planes = 1
    N = K = 256
    bufs = [(mk(M, K, planes), mk(M, N, planes)) for _ in range(3)]   # rotate: 3 x 393 MB >> L2
    B = mk(N, K, planes)
    bias = torch.randn(N, device=DEV)
    aux = mk(M, N, planes)
    aux.t.abs_()
    add = mk(M, N, planes)
    mask = torch.zeros(M, 32, dtype=torch.uint8, device=DEV)
    it = [0]

    def run(**kw):
        def f():
            A, O = bufs[it[0] % 3]
            it[0] += 1
            ops.linear(A, B, M, N, K, out=O, **kw)
        return f
    cases = {
        "plain": dict(),
        "bias+relu+mask_out": dict(bias=bias, act=1, mask_out=mask),
        "mask_in": dict(mask_in=mask),
        "bias+softplus": dict(bias=bias, act=2),
        "aux2": dict(aux=aux, aux_mode=2),
        "aux2+add": dict(aux=aux, aux_mode=2, add=add),
    }
    only = os.environ.get("CASE")
    for name, kw in cases.items():
        if only and only != name:
            continue
        ms = timeit(run(**kw))
        nbytes = 2.0 * M * (K + N) + (2.0 * M * N if "aux" in kw else 0) + (2.0 * M * N if "add" in kw else 0)
        print(f"linear {name:22s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s (algorithmic)  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
    if not only:
        zx = [(mk(M, 256, planes), mk(M, 256, planes)) for _ in range(3)]
        Z, X = zx[0]
        dW = torch.zeros(256, 256, device=DEV)
        db = torch.zeros(256, device=DEV)

        def dwrun(with_db):
            def f():
                Zi, Xi = zx[it[0] % 3]
                it[0] += 1
                ops.linear_dw(Zi, Xi, M, 256, 256, dW, db=db if with_db else None)
            return f
        for with_db in (False, True):
            ms = timeit(dwrun(with_db))
            print(f"dw 256x256 db={int(with_db)}        {ms*1e3:8.1f} us  {2.0*M*512/ms/1e6:7.0f} GB/s  "
                  f"{2.0*M*65536/ms/1e9:7.1f} TFLOP/s", flush=True)
        s = torch.zeros(256, device=DEV)
        ms = timeit(lambda: ops.colsum(Z, M, 256, s))
        print(f"colsum                 {ms*1e3:8.1f} us  {2.0*M*256/ms/1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
