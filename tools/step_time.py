"""Per-kernel timing of one SOM step through the production C-ABI entry points (run on a B200 through gpurun).

    python tools/step_time.py [B K_rows K_cols D [euclidean|cosine [tf32x3|fp16x3]]]

Every entry point is launched `iters` times back to back behind a GPU-side sleep (so the events see GPU time
only, L2-warm) and, for the CTA-pair GEMMs, once more with phase stamps.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vit_som_b200 import _lib, ops  # noqa: E402


def timed(name, fn, iters=20, stamps=False):
    L = _lib.lib()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    extra = ""
    if stamps:
        tb = torch.zeros(16, dtype=torch.int64, device="cuda")
        L.som_set_debug_times(tb.data_ptr())
        fn()
        torch.cuda.synchronize()
        L.som_set_debug_times(None)
        t = tb.cpu().tolist()
        names = ["setup", "prod_done", "mma_issued", "acc_ready", "epi_done", "synced", "freed", "acc_in_regs", "pre_epi"]
        extra = "  | " + " ".join(f"{n}={(t[i + 1] - t[0]) / 1e3:.1f}" for i, n in enumerate(names) if t[i + 1])
        if t[13]:     # epilogue cycle counters of warp 4 / CTA 0 (1.965 GHz): TMEM load, math, store per 16-column block
            extra += (f" || epi blocks={t[13]} ld={t[10] / 1965 / t[13]:.2f} math={t[11] / 1965 / t[13]:.2f} "
                      f"st={t[12] / 1965 / t[13]:.2f} us/block; sk wait={t[14] / 1965:.1f} add={t[15] / 1965:.1f} us")
    print(f"{name:28s} {us:9.1f} us{extra}", flush=True)
    return us


def main():
    a = sys.argv[1:]
    B, kr, kc, D = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (1024, 40, 40, 3136)
    fcn = a[4] if len(a) > 4 else "euclidean"
    prec = a[5] if len(a) > 5 else ops.DEFAULT_PRECISION
    mode = ops.MODE[fcn] | ops.PREC[prec]
    f16 = ops.is_f16(mode)
    K = kr * kc
    L = _lib.lib()
    sp = _lib.stream_ptr
    torch.manual_seed(0)
    x = torch.randn(B, D, device="cuda")
    W = torch.rand(K, D, device="cuda")
    pos = torch.stack([torch.arange(kr).repeat_interleave(kc), torch.arange(kc).repeat(kr)], 1).float().cuda()
    T = torch.full((1,), 10.0, device="cuda")
    g = torch.ones(1, device="cuda")
    xs, ws = ops.Staging(B, D, mode, x.device), ops.Staging(K, D, mode, x.device)
    ldd = (K + 3) // 4 * 4
    dist = torch.empty(B, ldd, device="cuda")
    packed = torch.empty(B, dtype=torch.int64, device="cuda")
    bmu = torch.empty(B, dtype=torch.int64, device="cuda")
    nrp, ncp = ops.loss_parts(B, K, x.device)                  # partial-sum tables of the loss kernel
    ldr = (K + 7) // 8 * 8 if f16 else ldd                     # fp16: R is a pair of half matrices
    r_floats = B * ldr // 2 if f16 else B * ldr
    row_floats = (B * nrp + 4 + 3) // 4 * 4
    rbuf = torch.empty(2 * r_floats + row_floats + ncp * K + 4, device="cuda")
    r_hi, r_lo = rbuf.data_ptr(), rbuf.data_ptr() + 4 * r_floats
    row_sum, col_sum = rbuf.data_ptr() + 8 * r_floats, rbuf.data_ptr() + 8 * r_floats + 4 * row_floats
    loss = torch.empty((), device="cuda")
    scratch = torch.zeros(1 << 16, device="cuda")
    dx, dw = torch.empty(B, D, device="cuda"), torch.empty(K, D, device="cuda")
    gws, gws_n = ops.gemm_workspace(x.device)
    SQUARE = os.environ.get("SOM_GENERAL_GRID") is None
    use_ws = os.environ.get("SOM_NO_WS") is None
    L.som_set_debug(int(os.environ.get("SOM_DEBUG", "0")))
    if os.environ.get("SOM_BN"):
        L.som_set_tuning(int(os.environ["SOM_BN"]), 0)          # force the tile width of every GEMM
    wsp, wsn = (gws, gws_n) if use_ws else (None, 0)

    def chk(rc, what):
        _lib.check(rc, what)

    total = 0.0
    total += timed("prep x+W (+packed reset)", lambda: chk(L.som_forward(
        x.data_ptr(), D, W.data_ptr(), D, B, K, D, mode, 1, 0, xs.hi, xs.lo, xs.aux, ws.hi, ws.lo, ws.aux, xs.ld,
        None, ldd, packed.data_ptr(), None, K, wsp, wsn, sp()), "fwd") if False else chk(L.som_prep_rows(
            x.data_ptr(), B, D, D, mode, xs.hi, xs.lo, xs.ld, xs.aux, sp()), "prep") or chk(L.som_prep_rows(
                W.data_ptr(), K, D, D, mode, ws.hi, ws.lo, ws.ld, ws.aux, sp()), "prep"))
    chk(L.som_bmu_init(packed.data_ptr(), B, sp()), "init")
    total += timed("GEMM fwd (dist + argmin)", lambda: chk(L.som_fwd_distances(
        xs.hi, xs.lo, xs.ld, xs.aux, ws.hi, ws.lo, ws.ld, ws.aux, B, K, D, mode, 0, dist.data_ptr(), ldd,
        packed.data_ptr(), wsp, wsn, sp()), "fwd"), stamps=True)
    timed("GEMM fwd (argmin only)", lambda: chk(L.som_fwd_distances(
        xs.hi, xs.lo, xs.ld, xs.aux, ws.hi, ws.lo, ws.ld, ws.aux, B, K, D, mode, 0, None, ldd,
        packed.data_ptr(), wsp, wsn, sp()), "fwd"), stamps=True)
    if f16:
        total += timed("decode + R scale statistic", lambda: chk(L.som_bmu_decode_scaled(
            packed.data_ptr(), B, K, bmu.data_ptr(), None, xs.aux, ws.aux, K, mode, sp()), "dec"))
    else:
        total += timed("decode", lambda: chk(L.som_bmu_decode(packed.data_ptr(), B, K, bmu.data_ptr(), None, sp()), "dec"))
    total += timed("loss + coeffs", lambda: chk(L.som_loss_fused(
        dist.data_ptr(), ldd, bmu.data_ptr(), pos.data_ptr(), kr if SQUARE else 0, kc if SQUARE else 0, B, K, 0,
        T.data_ptr(), 1.0 / (B * K), mode, r_hi, r_lo,
        ldr, row_sum, col_sum, scratch.data_ptr(), loss.data_ptr(), xs.aux if f16 else None, ws.aux if f16 else None,
        sp()), "loss"))
    total += timed("GEMM dW", lambda: chk(L.som_backward_dw(
        r_hi, r_lo, ldr, xs.hi, xs.lo, xs.ld, W.data_ptr(), D, col_sum, ncp, ws.aux, g.data_ptr(), B, K, D, mode,
        dw.data_ptr(), D, 0, 0, wsp, wsn, sp()), "dw"), stamps=True)
    total += timed("GEMM dx", lambda: chk(L.som_backward_dx(
        r_hi, r_lo, ldr, ws.hi, ws.lo, ws.ld, x.data_ptr(), D, row_sum, nrp, xs.aux, g.data_ptr(), B, K, D, mode,
        dx.data_ptr(), D, 0, 0, wsp, wsn, sp()), "dx"), stamps=True)
    timed("GEMM dW+dx fused launch", lambda: chk(L.som_backward_fused(
        r_hi, r_lo, ldr, xs.hi, xs.lo, ws.hi, ws.lo, xs.ld, x.data_ptr(), D, W.data_ptr(), D, row_sum, nrp, col_sum,
        ncp, xs.aux, ws.aux, g.data_ptr(), B, K, D, mode, dw.data_ptr(), D, 0, dx.data_ptr(), D, 0, 0, None, None,
        wsp, wsn, sp()), "bwd"), stamps=True)
    # the fused prototype optimizer step (update + staging): 9 arrays of K x D floats through HBM
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    hp = torch.tensor([1e-3, 1.0, 1.0], device="cuda")
    t_adam = timed("AdamW + W staging", lambda: chk(L.som_adamw_step(
        W.data_ptr(), D, dw.data_ptr(), D, m.data_ptr(), v.data_ptr(), D, K, D, hp.data_ptr(), 0.9, 0.999, 1e-8, 0.01,
        mode, ws.hi, ws.lo, ws.ld, ws.aux, sp()), "adamw"))
    print(f"   AdamW pass: {9 * K * D * 4 / t_adam / 1e3:.0f} GB/s (9 K D floats; L2-warm)")
    per = 8.0 if f16 else 12.0
    print(f"   (loss kernel moves {per:.0f} B per element: {per * B * K / 1e6:.1f} MB)")
    print(f"sum of step kernels: {total:.1f} us  (B={B} K={K} D={D} {fcn} {prec}, workspace={'yes' if use_ws else 'no'})")


if __name__ == "__main__":
    main()
