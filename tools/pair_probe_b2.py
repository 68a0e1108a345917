"""Cold-weight probe of the CFG-batch-2 conv3x3 / deep-K GEMM shapes: single CTAs vs forced CTA pairs (cta_group::2), tile widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd._lib import lib
DEV = "cuda:0"
L = lib()
from tools.cold_probe import timeit  # noqa

def case(M, N, K, conv, variants):
    C = K // 9 if conv else K
    ncopy = min(400, max(8, int(300e6 / (N * K * 2)) + 1))
    a = torch.randn(M, C, device=DEV).bfloat16()
    ws = [(torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16() for _ in range(ncopy)]
    bias = torch.randn(N, device=DEV)
    out = torch.randn(M, N, device=DEV)
    row = f"{'conv' if conv else 'gemm'} M{M} N{N} K{K}:"
    for pair, bn in variants:
        try:
            argl = [ops.gemm(a, w, out, bias=bias, residual=out, conv=conv, block_n=bn, pair=pair, launch=False) for w in ws]
            fns = [(lambda x=x: ops.gemm_run(x)) for x in argl]
            row += f" [pair{pair} bn{bn}] {timeit(fns):5.1f}"
        except Exception as e:
            row += f" [pair{pair} bn{bn}] err {str(e)[:30]}"
    print(row, flush=True)

V = ((0, 0), (1, 0), (1, 160), (1, 320), (-1, 160), (-1, 320))
case(8192, 320, 2880, (2, 64, 64), ((0, 0), (1, 0), (1, 160), (-1, 160)))
case(8192, 320, 5760, (2, 64, 64), ((0, 0), (1, 0), (1, 160), (-1, 160)))
case(8192, 320, 8640, (2, 64, 64), ((0, 0), (1, 0), (1, 160), (-1, 160)))
case(8192, 640, 5760, (2, 64, 64), ((0, 0), (1, 0), (1, 160), (1, 128), (-1, 160)))
case(2048, 640, 5760, (2, 32, 32), ((0, 0), (1, 0), (1, 160), (1, 128), (1, 64)))
case(2048, 640, 11520, (2, 32, 32), ((0, 0), (1, 0), (1, 160), (1, 128), (1, 64)))
case(2048, 640, 17280, (2, 32, 32), ((0, 0), (1, 0), (1, 160), (1, 128), (1, 64)))
case(2048, 1280, 11520, (2, 32, 32), ((0, 0), (1, 0), (1, 160), (1, 128)))
