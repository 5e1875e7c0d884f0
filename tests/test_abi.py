"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/vit_unet_b200.h, and the host layer refuses to run without CUDA (no silent fallback)."""
import contextlib
import io
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from vit_unet_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, "include", "vit_unet_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(vu_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 25
    lib = built.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    bound = set(built.SIGNATURES) | set(built._SPECIAL)
    assert declared == bound, declared ^ bound


def test_abi_version_and_error_text(built):
    lib = built.load()
    assert lib.vu_version() == built.ABI_VERSION
    # argument validation happens before any CUDA call -> safe without a GPU
    rc = lib.vu_repatch(None, None, 1, 3, 32, 32, 0, 16, None)
    assert rc == 1 and b"vu_repatch" in lib.vu_last_error()
    rc = lib.vu_softmax_rows(None, 0, 0, 0, 1.0, None)
    assert rc == 1


def test_gemm_desc_layout_matches_header(built):
    import ctypes as C
    # field order/size of the ctypes mirror vs the C struct: compile a probe with gcc and compare sizeof/offsets
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "vit_unet_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(vu_gemm_desc), offsetof(vu_gemm_desc, M),
      offsetof(vu_gemm_desc, lda), offsetof(vu_gemm_desc, sAo), offsetof(vu_gemm_desc, alpha),
      offsetof(vu_gemm_desc, precision)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).split()
    G = built.GemmDesc
    exp = [C.sizeof(G), G.M.offset, G.lda.offset, G.sAo.offset, G.alpha.offset, G.precision.offset]
    assert [int(v) for v in out] == exp


def test_no_cpu_fallback(built):
    import vit_unet_b200 as vu
    from vit_unet_b200 import ops
    with contextlib.redirect_stdout(io.StringIO()):
        net = vu.get_vit_unet("lite")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 224, 224))
    with pytest.raises(ops.VuError):
        ops.repatch(torch.zeros(4), torch.zeros(4), 1, 1, 2, 2, 0, 2)
    with pytest.raises(ValueError):
        vu.get_vit_unet("nope")


def test_module_surface_matches_reference_contract(built):
    """state_dict keys / shapes / parameter counts of the drop-in modules."""
    import vit_unet_b200 as vu
    from oracle import vit_unet_oracle as O
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        net = vu.get_vit_unet("base")
        ref = O.get_vit_unet("base")
    assert "Architecture information:" in buf.getvalue()          # model.py:301-307 prints at construction
    sa, sb = net.state_dict(), ref.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape for k in sa)
    assert sum(p.numel() for p in net.parameters()) == 39_623_512
    with contextlib.redirect_stdout(io.StringIO()):
        rd = vu.ViT_UNet(depth=2, depth_te=2, size_bottleneck=2, preprocessing='conv', num_patches=49,
                         patch_size=32, num_channels=3, hidden_dim=128, num_heads=8, attn_drop=.2, proj_drop=.2,
                         linear_drop=0, dtype=torch.float32)
    assert sum(p.numel() for p in rd.parameters()) == 36_613_036   # README.md:34
    order = net._param_names
    assert order[0].startswith("PE.") and order[-1].startswith("conv2d.")
    assert sorted(order) == sorted(n for n, _ in net.named_parameters())
