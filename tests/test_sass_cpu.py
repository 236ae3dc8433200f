"""The built library really is Blackwell-native: SASS of libsom_b200.so (cuobjdump, no GPU needed) holds the tcgen05 /
TMA / TMEM instructions in every instantiation of the GEMM kernels - both operand precisions - and the packed fp16
conversions in the fp16 staging kernels.  (profiles/*_sass.txt is the same listing, committed as evidence.)"""
import re
import shutil
import subprocess

import pytest

from vit_som_b200 import _lib


@pytest.fixture(scope="module")
def sass():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.which(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.build()
    out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per_fn, fn = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = []
        elif fn is not None:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m:
                per_fn[fn].append(m.group(1))
    return per_fn


def _kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_every_gemm_instantiation_uses_tcgen05_tma_and_tmem(sass):
    pair = _kernels(sass, "som_gemm3x_pair_kernel")
    single = _kernels(sass, "som_gemm3x_kernelI")
    assert len(pair) == 6 and len(single) == 6            # 3 epilogues x 2 operand precisions each
    for name, ops in {**pair, **single}.items():
        assert sum(o.startswith("UTCHMMA") for o in ops) == 12, name          # 4 k-steps x 3 products per k-block
        assert any(o.startswith("UTMALDG") for o in ops), name                # TMA loads
        assert any(o.startswith("LDTM") for o in ops), name                   # tcgen05.ld
        assert not any(o.startswith("HMMA") or o.startswith("HGMMA") for o in ops), name   # no legacy mma.sync / wgmma
    for name, ops in pair.items():
        assert all(o.startswith("UTCHMMA.2CTA") for o in ops if o.startswith("UTCHMMA")), name      # cta_group::2
        assert any(o.startswith("UTMALDG.3D") for o in ops), name             # MN-major tiles: one 3-D operation per tile
        assert any(o.startswith("STTM") for o in ops), name                   # totals written back to TMEM for the epilogue loop


def test_fp16_staging_kernels_use_packed_conversions(sass):
    for needle in ("prep_rows_kernelILi256ELb1", "prep_rows_kernelILi32ELb1", "adamw_stage_kernelILi256ELb1",
                   "adamw_stage_kernelILi32ELb1", "loss_coeffs_fast_kernelILi0ELb1", "loss_coeffs_fast_kernelILi1ELb1"):
        ks = _kernels(sass, needle)
        assert len(ks) == 1, needle
        ops = next(iter(ks.values()))
        assert any(o.startswith("F2FP") for o in ops), needle                  # cvt.rn.f16x2.f32
    for needle in ("prep_rows_kernelILi256ELb0", "loss_coeffs_fast_kernelILi0ELb0"):
        ops = next(iter(_kernels(sass, needle).values()))
        assert not any(o.startswith("F2FP") for o in ops), needle              # the tf32 instantiations stay fp32


def test_multi_gpu_exchange_kernel_uses_multimem(sass):
    ops = next(iter(_kernels(sass, "nvls_allreduce_mean_kernel").values()))
    assert any("LDGMC" in o or "MULTIMEM" in o for o in ops)                  # multimem.ld_reduce
