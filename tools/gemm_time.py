"""Timing experiments for the tcgen05 GEMM mainloop (run on a B200 through gpurun).

    python tools/gemm_time.py "M,N,K,a_mn,b_mn,cg,bn,kchunk,debug[,sk]" [...more specs]

sk: stream-K policy (-1 never / no workspace, 0 cost model, 1 whenever possible); default -1.

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
        fields = [int(t) for t in spec.split(",")]
        M, N, K, a_mn, b_mn, cg, bn, kchunk, debug = fields[:9]
        sk = fields[9] if len(fields) > 9 else -1
        WS = torch.zeros(int(L.som_gemm_workspace_floats()), device="cuda") if sk >= 0 else None
        L.som_set_streamk(sk)
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
                                  b_mn, M, N, K, bn, kchunk, 3, C.data_ptr(), N, WS.data_ptr() if WS is not None else None, WS.numel() if WS is not None else 0, s)
            _lib.check(rc, "som_debug_gemm")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        torch.cuda._sleep(20_000_000)      # ~10 ms: the launches below queue up, so the events time GPU work only
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        tf = 2.0 * M * N * K / us / 1e6
        stamps = ""
        if cg == 2 and os.environ.get("SOM_STAMPS"):
            tb = torch.zeros(8, dtype=torch.int64, device="cuda")
            L.som_set_debug_times(tb.data_ptr())
            run()
            torch.cuda.synchronize()
            L.som_set_debug_times(None)
            t = tb.cpu().tolist()
            names = ["setup", "prod_done", "mma_issued", "acc_ready", "epi_done", "synced", "freed"]
            stamps = "  | " + " ".join(f"{n}={(t[i + 1] - t[0]) / 1e3:.1f}" for i, n in enumerate(names) if t[i + 1])
        print(f"{spec:45s} {us:9.1f} us  {tf:7.1f} TF(alg){stamps}", flush=True)
    L.som_set_debug(0)
    L.som_set_cta_group(0)
    L.som_set_streamk(0)


if __name__ == "__main__":
    main()
