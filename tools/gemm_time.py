"""Timing experiments for the tcgen05 GEMM mainloop (run on a B200 through gpurun).

    python tools/gemm_time.py "M,N,K,a_mn,b_mn,cg,bn,kchunk,debug" [...more specs]

Each spec launches the diagnostic GEMM 20 times back to back (operands L2-warm) and prints the mean time.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vit_som_b200 import _lib  # noqa: E402


def main():
    L = _lib.lib()
    s = _lib.stream_ptr()
    for spec in sys.argv[1:]:
        M, N, K, a_mn, b_mn, cg, bn, kchunk, debug = [int(t) for t in spec.split(",")]
        torch.manual_seed(0)
        lda = (M if a_mn else K)
        ldb = (N if b_mn else K)
        lda, ldb = (lda + 3) // 4 * 4, (ldb + 3) // 4 * 4
        a_hi = torch.randn((K if a_mn else M), lda, device="cuda")
        a_lo = a_hi * 1e-4
        b_hi = torch.rand((K if b_mn else N), ldb, device="cuda")
        b_lo = b_hi * 1e-4
        C = torch.empty(M, N, device="cuda")
        L.som_set_cta_group(cg)
        L.som_set_debug(debug)

        def run():
            rc = L.som_debug_gemm(a_hi.data_ptr(), a_lo.data_ptr(), lda, a_mn, b_hi.data_ptr(), b_lo.data_ptr(), ldb,
                                  b_mn, M, N, K, bn, kchunk, 3, C.data_ptr(), N, s)
            _lib.check(rc, "som_debug_gemm")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        tf = 2.0 * M * N * K / us / 1e6
        print(f"{spec:45s} {us:9.1f} us  {tf:7.1f} TF(alg)", flush=True)
    L.som_set_debug(0)
    L.som_set_cta_group(0)


if __name__ == "__main__":
    main()
