"""GPU bring-up probe for the tcgen05 mainloop (run on a B200 through gpurun).

Each case runs in its own subprocess with a timeout so that a trap or a protocol bug in one case
cannot take the others down.  Results go to gpurun_out/probe.json.

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py --case kk_exact
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


WS = None      # GEMM workspace (stream-K); allocated when SOM_PROBE_SK is set


def _setup_streamk(L):
    global WS
    import torch
    mode = os.environ.get("SOM_PROBE_SK")
    if mode is None:
        return
    L.som_set_streamk(int(mode))
    if WS is None and int(mode) >= 0:
        WS = torch.zeros(int(L.som_gemm_workspace_floats()), device="cuda")


def split_tf32(x):
    """Exact tf32 hi/lo split with round-to-nearest-away (cvt.rna.tf32.f32)."""
    import torch
    def rna(v):
        i = v.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    hi = rna(x)
    lo = rna(x - hi)
    return hi, lo


def run_gemm(A, Bm, a_mn, b_mn, bn=0, kchunk=0, passes=3):
    """A: [M,Kr], Bm: [N,Kr] logical. Stores them in the requested major and calls the debug GEMM."""
    import torch
    from vit_som_b200 import _lib
    L = _lib.lib()
    L.som_set_cta_group(int(os.environ.get("SOM_PROBE_CG", "0")))
    L.som_set_debug(int(os.environ.get("SOM_PROBE_DEBUG", "0")))
    _setup_streamk(L)
    M, Kr = A.shape
    N = Bm.shape[0]

    def stage(X, mn):
        rows, cols = (X.shape[1], X.shape[0]) if mn else X.shape
        ld = (cols + 3) // 4 * 4
        buf_hi = torch.zeros(rows, ld, device="cuda")
        buf_lo = torch.zeros(rows, ld, device="cuda")
        src = X.t().contiguous() if mn else X
        if passes == 3:
            hi, lo = split_tf32(src)
        else:
            hi, lo = src, torch.zeros_like(src)
        buf_hi[:, :cols] = hi
        buf_lo[:, :cols] = lo
        return buf_hi, buf_lo, ld

    a_hi, a_lo, lda = stage(A, a_mn)
    b_hi, b_lo, ldb = stage(Bm, b_mn)
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = L.som_debug_gemm(a_hi.data_ptr(), a_lo.data_ptr(), lda, a_mn, b_hi.data_ptr(), b_lo.data_ptr(), ldb, b_mn,
                          M, N, Kr, bn, kchunk, passes, C.data_ptr(), N, WS.data_ptr() if WS is not None else None, WS.numel() if WS is not None else 0, _lib.stream_ptr())
    _lib.check(rc, "som_debug_gemm")
    torch.cuda.synchronize()
    return C


def err_stats(C, ref):
    import torch
    d = (C.double() - ref)
    return {
        "max_abs": d.abs().max().item(),
        "rel_fro": (d.norm() / ref.norm()).item(),
        "mean_signed_rel": (d / ref.abs().clamp_min(1e-30)).mean().item(),
        "nan": int(torch.isnan(C).sum().item()),
    }


def case_exact(a_mn, b_mn, M=128, N=128, Kr=64, bn=0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randint(-4, 5, (M, Kr), generator=g).float().cuda()
    Bm = torch.randint(-4, 5, (N, Kr), generator=g).float().cuda()
    C = run_gemm(A, Bm, a_mn, b_mn, bn=bn, passes=1)
    ref = A.double() @ Bm.double().t()
    st = err_stats(C, ref)
    st["ok"] = st["max_abs"] == 0.0 and st["nan"] == 0
    if not st["ok"]:
        bad = (C.double() != ref).nonzero()
        st["first_bad"] = bad[:8].tolist()
        st["n_bad"] = int(bad.shape[0])
        st["C00"] = C[:2, :4].tolist()
        st["ref00"] = ref[:2, :4].tolist()
    return st


def case_random(a_mn, b_mn, M, N, Kr, bn=0, kchunk=0, passes=3, positive=False):
    import torch
    g = torch.Generator(device="cpu").manual_seed(2)
    A = (torch.rand(M, Kr, generator=g) if positive else torch.randn(M, Kr, generator=g)).cuda()
    Bm = torch.rand(N, Kr, generator=g).cuda()
    if passes == 1:   # make the inputs exact tf32 so that only accumulation error remains
        A, _ = split_tf32(A)
        Bm, _ = split_tf32(Bm)
    C = run_gemm(A, Bm, a_mn, b_mn, bn=bn, kchunk=kchunk, passes=passes)
    ref = A.double() @ Bm.double().t()
    st = err_stats(C, ref)
    ref32 = (A @ Bm.t())
    torch.backends.cuda.matmul.allow_tf32 = False
    st["torch_fp32_rel_fro"] = ((ref32.double() - ref).norm() / ref.norm()).item()
    st["ok"] = st["nan"] == 0 and st["rel_fro"] < (1e-5 if passes == 3 else 1e-3)
    return st


def case_timing():
    import torch
    out = {}
    g = torch.Generator(device="cpu").manual_seed(3)
    for name, (M, N, Kr, a_mn, b_mn) in {
        "fwd_cfg2": (1024, 1600, 3136, 0, 0),
        "dx_cfg2": (1024, 3136, 1600, 0, 1),
        "dw_cfg2": (1600, 3136, 1024, 1, 1),
    }.items():
        A = torch.randn(M, Kr, generator=g).cuda()
        Bm = torch.rand(N, Kr, generator=g).cuda()
        for bn in [int(t) for t in os.environ.get("SOM_PROBE_BNS", "0,128,64").split(",")]:
            C = run_gemm(A, Bm, a_mn, b_mn, bn=bn)
            ref = A.double() @ Bm.double().t()
            st = err_stats(C, ref)
            # timing: re-stage once, launch many
            from vit_som_b200 import _lib
            L = _lib.lib()
            _setup_streamk(L)

            def stage(X, mn):
                src = X.t().contiguous() if mn else X
                hi, lo = split_tf32(src)
                return hi.contiguous(), lo.contiguous(), src.shape[1]
            a_hi, a_lo, lda = stage(A, a_mn)
            b_hi, b_lo, ldb = stage(Bm, b_mn)
            Cb = torch.empty(M, N, device="cuda")
            s = _lib.stream_ptr()
            for _ in range(3):
                L.som_debug_gemm(a_hi.data_ptr(), a_lo.data_ptr(), lda, a_mn, b_hi.data_ptr(), b_lo.data_ptr(), ldb,
                                 b_mn, M, N, Kr, bn, 0, 3, Cb.data_ptr(), N, WS.data_ptr() if WS is not None else None, WS.numel() if WS is not None else 0, s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            iters = 20
            for _ in range(iters):
                L.som_debug_gemm(a_hi.data_ptr(), a_lo.data_ptr(), lda, a_mn, b_hi.data_ptr(), b_lo.data_ptr(), ldb,
                                 b_mn, M, N, Kr, bn, 0, 3, Cb.data_ptr(), N, WS.data_ptr() if WS is not None else None, WS.numel() if WS is not None else 0, s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            st["ms"] = ms
            try:
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByIndex(0)
                st["sm_mhz_after"] = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            except Exception:  # noqa: BLE001
                pass
            st["tflops_alg"] = 2.0 * M * N * Kr / ms / 1e9
            out[f"{name}_bn{bn}"] = st
    out["ok"] = True
    return out


CASES = {
    "kk_exact": lambda: case_exact(0, 0),
    "kk_exact_bn64": lambda: case_exact(0, 0, N=256, bn=64),
    "kk_exact_multi": lambda: case_exact(0, 0, M=384, N=512, Kr=256),
    "kk_exact_ragged": lambda: case_exact(0, 0, M=200, N=176, Kr=100),
    "kmn_exact": lambda: case_exact(0, 1),
    "mnk_exact": lambda: case_exact(1, 0),
    "mnmn_exact": lambda: case_exact(1, 1),
    "kmn_exact_ragged": lambda: case_exact(0, 1, M=200, N=176, Kr=100),
    "mnmn_exact_ragged": lambda: case_exact(1, 1, M=200, N=176, Kr=100),
    "kk_exact_n16": lambda: case_exact(0, 0, M=128, N=16, Kr=512),
    "kk_rand3": lambda: case_random(0, 0, 256, 256, 3136),
    "kk_rand3_nochunk": lambda: case_random(0, 0, 256, 256, 3136, kchunk=4096),
    "kk_rand1_pos_nochunk": lambda: case_random(0, 0, 128, 128, 4096, kchunk=4096, passes=1, positive=True),
    "kk_rand1_pos_chunk16": lambda: case_random(0, 0, 128, 128, 4096, kchunk=16, passes=1, positive=True),
    "kk_rand1_pos_chunk4": lambda: case_random(0, 0, 128, 128, 4096, kchunk=4, passes=1, positive=True),
    "kk_rand3_pos_nochunk": lambda: case_random(0, 0, 128, 128, 4096, kchunk=4096, passes=3, positive=True),
    "kk_rand3_pos_chunk16": lambda: case_random(0, 0, 128, 128, 4096, kchunk=16, passes=3, positive=True),
    "kk_rand3_long": lambda: case_random(0, 0, 512, 1600, 49152),
    "mnmn_rand3": lambda: case_random(1, 1, 1600, 3136, 1024),
    "kmn_rand3": lambda: case_random(0, 1, 1024, 3136, 1600),
    "timing": case_timing,
    # shapes that exercise the CTA-pair kernel (M > 128): run with --cg 2
    "p_kk_exact_256": lambda: case_exact(0, 0, M=256, N=256, Kr=64, bn=256),
    "p_kk_exact_128": lambda: case_exact(0, 0, M=256, N=256, Kr=64, bn=128),
    "p_kk_exact_64": lambda: case_exact(0, 0, M=256, N=256, Kr=64, bn=64),
    "p_kk_exact_multi": lambda: case_exact(0, 0, M=768, N=1024, Kr=512, bn=256),
    "p_kk_exact_ragged": lambda: case_exact(0, 0, M=300, N=336, Kr=100),
    "p_kmn_exact": lambda: case_exact(0, 1, M=256, N=256, Kr=64, bn=256),
    "p_mnk_exact": lambda: case_exact(1, 0, M=256, N=256, Kr=64, bn=256),
    "p_mnmn_exact": lambda: case_exact(1, 1, M=256, N=256, Kr=64, bn=256),
    "p_mnmn_exact_ragged": lambda: case_exact(1, 1, M=300, N=336, Kr=100),
    "p_kk_rand3_256": lambda: case_random(0, 0, 1024, 1600, 3136, bn=256),
    "p_kk_rand3_192": lambda: case_random(0, 0, 1024, 1600, 3136, bn=192),
    "p_kk_rand3_pos": lambda: case_random(0, 0, 512, 512, 4096, positive=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated substrings: run the cases containing any")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.json"))
    ap.add_argument("--cg", type=int, default=None, help="force the kernel: 1 single-CTA, 2 CTA pair (default: cost model)")
    ap.add_argument("--debug", type=int, default=0, help="GemmShape.debug bits (1: no TMA after prologue, 2: no MMA)")
    ap.add_argument("--sk", type=int, default=None, help="stream-K policy with a workspace: -1 never, 0 cost model, 1 always")
    ap.add_argument("--bns", default=None, help="tile widths for the timing case, e.g. 0,256,128")
    args = ap.parse_args()
    if args.cg is not None:
        os.environ["SOM_PROBE_CG"] = str(args.cg)
    if args.sk is not None:
        os.environ["SOM_PROBE_SK"] = str(args.sk)
    if args.debug:
        os.environ["SOM_PROBE_DEBUG"] = str(args.debug)
    if args.bns is not None:
        os.environ["SOM_PROBE_BNS"] = args.bns
    if args.case:
        try:
            res = CASES[args.case]()
        except Exception as exc:  # noqa: BLE001
            res = {"ok": False, "error": repr(exc)}
        print("PROBE_RESULT " + json.dumps(res))
        return
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    results = {}
    for name in CASES:
        if args.only and not any(tok in name for tok in args.only.split(",")):
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=180)
            line = [l for l in p.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
            if line:
                results[name] = json.loads(line[-1][len("PROBE_RESULT "):])
            else:
                results[name] = {"ok": False, "rc": p.returncode, "stdout": p.stdout[-2000:], "stderr": p.stderr[-2000:]}
        except subprocess.TimeoutExpired:
            results[name] = {"ok": False, "error": "timeout"}
        results[name]["secs"] = round(time.time() - t0, 1)
        print(name, json.dumps(results[name])[:600], flush=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
