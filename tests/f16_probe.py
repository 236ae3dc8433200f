"""GPU probe of the 3xFP16 path (run through gpurun; lives under tests/ because it uses the oracle as its checker): GEMM mainloop in every operand layout against fp64, accuracy next
to 3xTF32 on the same data, timings of the three GEMM shapes of config 2 / a config-5 chunk, then the layer end to end
(forward, loss, backward) against the fp64 oracle in both precisions."""
import ctypes
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_som_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
if os.environ.get("SOM_KCHUNK"):
    L.som_set_tuning(0, int(os.environ["SOM_KCHUNK"]))     # k-blocks per accumulation chain (the pair kernel uses half)
dev = torch.device("cuda:0")
sp = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def split_tf32(a):
    hi = (a.view(torch.int32) + 0x1000 & ~0x1fff).view(torch.float32)
    lo = a - hi
    lo = (lo.view(torch.int32) + 0x1000 & ~0x1fff).view(torch.float32)
    return hi.contiguous(), lo.contiguous()


def split_f16(a):
    hi = a.half()
    lo = (a - hi.float()).half()
    return hi.contiguous(), lo.contiguous()


def pad_cols(t, mult):
    r, c = t.shape
    cp = (c + mult - 1) // mult * mult
    if cp == c:
        return t.contiguous()
    out = torch.zeros((r, cp), device=t.device, dtype=t.dtype)
    out[:, :c] = t
    return out


def run_gemm(A, B, a_mn, b_mn, f16, bn=0, cg=0, passes=3, reps=0):
    """A [M, Kr], B [N, Kr] fp32 logical operands; returns C [M, N] (and ms per call when reps > 0)."""
    M, Kr = A.shape
    N = B.shape[0]
    mult = 8 if f16 else 4
    sa = pad_cols(A.t().contiguous() if a_mn else A, mult)
    sb = pad_cols(B.t().contiguous() if b_mn else B, mult)
    (a_hi, a_lo), (b_hi, b_lo) = (split_f16(sa), split_f16(sb)) if f16 else (split_tf32(sa), split_tf32(sb))
    C = torch.full((M, (N + 3) // 4 * 4), float("nan"), device=dev)
    ws, wsn = ops.gemm_workspace(dev)
    L.som_set_cta_group(cg)
    fn = L.som_debug_gemm_f16 if f16 else L.som_debug_gemm
    call = lambda: fn(a_hi.data_ptr(), a_lo.data_ptr(), sa.shape[1], a_mn, b_hi.data_ptr(), b_lo.data_ptr(), sb.shape[1],  # noqa: E731
                      b_mn, M, N, Kr, bn, 0, passes, C.data_ptr(), C.shape[1], ws, wsn, sp())
    rc = call()
    if rc != 0:
        L.som_set_cta_group(0)
        raise RuntimeError(f"rc={rc}: {L.som_last_error().decode()}")
    torch.cuda.synchronize()
    ms = None
    if reps:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            call()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
    L.som_set_cta_group(0)
    return C[:, :N], ms


def rel(c, ref):
    return float((c.double() - ref).abs().max() / ref.abs().max())


def gemm_checks():
    torch.manual_seed(1)
    print("== GEMM layouts (rel. max error vs fp64; 3xTF32 | 3xFP16) ==")
    for (M, N, Kr) in [(512, 384, 320), (300, 200, 136), (256, 1600, 3136)]:
        A = torch.randn(M, Kr, device=dev)
        B = torch.rand(N, Kr, device=dev)
        ref = A.double() @ B.double().t()
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                for cg in (1, 2):
                    row = f"M={M} N={N} K={Kr} a_mn={a_mn} b_mn={b_mn} cg={cg}:"
                    for f16 in (0, 1):
                        try:
                            C, _ = run_gemm(A, B, a_mn, b_mn, f16, cg=cg)
                            row += f"  {rel(C, ref):.2e}"
                        except Exception as exc:  # noqa: BLE001
                            row += f"  FAIL({str(exc)[:80]})"
                    print(row, flush=True)
    # all-positive data: accumulator rounding bias
    A = torch.rand(512, 3136, device=dev) + 0.5
    B = torch.rand(768, 3136, device=dev) + 0.5
    ref = A.double() @ B.double().t()
    for f16 in (0, 1):
        C, _ = run_gemm(A, B, 0, 0, f16, cg=2)
        print(f"all-positive K=3136 {'fp16' if f16 else 'tf32'}: {rel(C, ref):.2e}  (torch fp32: {rel(A @ B.t(), ref):.2e})")


def gemm_timings():
    print("== GEMM timings (ms; EPI_RAW, 3 passes) ==")
    shapes = {"cfg2 fwd  (x.W^T)": (1024, 1600, 3136, 0, 0), "cfg2 dW   (R^T.x)": (1600, 3136, 1024, 1, 1),
              "cfg2 dx   (R.W)": (1024, 3136, 1600, 0, 1), "cfg5 fwd chunk": (8192, 16384, 256, 0, 0),
              "cfg5 dW chunk": (16384, 256, 8192, 1, 1), "cfg5 dx chunk": (8192, 256, 16384, 0, 1)}
    for name, (M, N, Kr, a_mn, b_mn) in shapes.items():
        A = torch.randn(M, Kr, device=dev)
        B = torch.randn(N, Kr, device=dev)
        row = f"{name:20s}"
        for f16 in (0, 1):
            try:
                _, ms = run_gemm(A, B, a_mn, b_mn, f16, reps=20)
                row += f"  {'fp16' if f16 else 'tf32'} {ms * 1e3:8.1f} us ({2 * M * N * Kr / ms / 1e9:6.1f} TF/s)"
            except Exception as exc:  # noqa: BLE001
                row += f"  FAIL({str(exc)[:80]})"
        print(row, flush=True)


def layer_checks():
    from oracle import som_oracle as O
    from oracle.ref_import import make_config
    from vit_som_b200 import SOMLayer
    print("== layer parity vs fp64 oracle (bmu mismatches / not near-tie, loss, dx, dW rel. errors) ==")
    cases = [([12, 12], 320, 200, 6.0), ([24, 24], 3136, 256, 10.0), ([40, 40], 3136, 1024, 20.0), ([7, 9], 50, 33, 2.0),
             ([40, 40], 3136, 1024, 0.5)]
    for fcn in ("euclidean", "cosine"):
        for (ms, D, Bn, T) in cases:
            for prec in ("tf32x3", "fp16x3"):
                try:
                    torch.manual_seed(0)
                    layer = SOMLayer(make_config(ms, D, fcn, Tmax=T, Tmin=0.1)).to(dev)
                    layer.precision = prec
                    x = torch.randn(Bn, D, device=dev, requires_grad=True)
                    d, bmu = layer(x)
                    w = layer.compute_weights(bmu)
                    loss = layer.som_loss(w, d)
                    loss.backward()
                    torch.cuda.synchronize()
                    xn, Wn = x.detach().cpu().numpy(), layer.prototypes.detach().cpu().numpy()
                    ref = O.step(xn, Wn, O.grid_positions(ms), T, fcn, 1.0, np.float64)
                    n_bad, hard, worst = O.classify_bmu_mismatches(xn, Wn, bmu.cpu().numpy(), fcn)
                    if n_bad:
                        ref = O.step(xn, Wn, O.grid_positions(ms), T, fcn, 1.0, np.float64, bmu_override=bmu.cpu().numpy())
                    print(f"{fcn:9s} map={ms} D={D} B={Bn} T={T} {prec}: bmu {n_bad}/{hard}  "
                          f"dist {O.rel_err(d.detach().cpu().numpy(), ref.distances):.2e}  "
                          f"loss {abs(loss.item() - float(ref.loss)) / abs(float(ref.loss)):.2e}  "
                          f"dx {O.rel_err(x.grad.cpu().numpy(), ref.grad_x):.2e}  "
                          f"dW {O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w):.2e}", flush=True)
                except Exception as exc:  # noqa: BLE001
                    print(f"{fcn} map={ms} D={D} B={Bn} {prec}: FAIL {exc!r}"[:300], flush=True)
                    traceback.print_exc()


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "time", "layer"]
    t0 = time.time()
    for w, fn in (("gemm", gemm_checks), ("time", gemm_timings), ("layer", layer_checks)):
        if w in which:
            try:
                fn()
            except Exception:  # noqa: BLE001
                traceback.print_exc()
            torch.cuda.synchronize()
    print(f"probe done in {time.time() - t0:.1f} s")
