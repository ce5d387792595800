"""CPU: the C-ABI library loads and exports every symbol include/sdrm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sdrm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdrm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    return ge.build()


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdrm_b200.h but not exported"


def test_ctypes_table_covers_header(built_lib):
    from sdrm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.sdrm_version() >= 100
    assert isinstance(lib.sdrm_last_error(), bytes)


def test_argument_validation_without_gpu(built_lib):
    """Pure host-side checks return error codes + messages; nothing is launched."""
    from sdrm_b200 import _lib
    lib = _lib.load()
    assert lib.sdrm_sample_workspace_bytes(None, 10) == 0
    assert lib.sdrm_probe_linear_workspace_bytes(128, 64, 64) > 0
    assert lib.sdrm_probe_linear_workspace_bytes(0, 64, 64) == 0
    rc = lib.sdrm_topk(None, 4, 10, 10, 5, None, None, None)
    assert rc == -1 and b"null" in lib.sdrm_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.sdrm_topk(p, 4, 10, 10, 65, p, None, None) == -2      # k > 64 unsupported
    assert lib.sdrm_topk(p, 4, 10, 5, 5, p, None, None) == -1        # ld < n_items
    assert lib.sdrm_sample(None, 1, 0, None, None, 0, None, None, 0, None, None, None, None, 0, None) == -1


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sdrm_b200 import _lib
    from sdrm_b200.models import SDRM, VAE
    from sdrm_b200.train_SDRM import sample_ddpm
    diff, vae = SDRM(8, 5, 8, 1), VAE(12, 6, 8)
    with pytest.raises(_lib.SdrmError):
        sample_ddpm(4, diff, vae, 8, 1.0, n_timesteps=5)
    with pytest.raises(_lib.SdrmError):
        from sdrm_b200 import metrics
        metrics.recall_at_k_batch(torch.zeros(2, 5).numpy(), torch.zeros(2, 5).numpy(), k=2)


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import sdrm_b200.train_SDRM, sdrm_b200.metrics, sdrm_b200.training, "
            "sdrm_b200.distributed, sdrm_b200.evaluators; assert not any(m.startswith('oracle') for m in sys.modules)" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_layer_geometry_normal_and_column_split(built_lib):
    """Host-only query of the chunk geometry (engine_host.cu make_geom / make_geom_split): chunks are multiples of 16 columns, at
    most 256 wide, cover N; the column-split geometry has at least 8 chunks (one per CTA of a cluster of 8) and never fewer than the
    normal one; K is cut into 64-wide blocks either way."""
    import ctypes as C
    lib = C.CDLL(built_lib)
    def geom(N, K, split):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        assert lib.sdrm_layer_geometry(N, K, split, C.byref(a), C.byref(b), C.byref(c)) == 0
        return a.value, b.value, c.value
    assert geom(830, 830, 0) == (4, 208, 13) and geom(830, 830, 1) == (8, 112, 13)      # cfg 1 denoiser layers
    assert geom(950, 950, 0) == (4, 240, 15) and geom(950, 950, 1) == (8, 128, 15)      # cfg 5
    assert geom(340, 340, 0) == (2, 176, 6) and geom(400, 400, 1) == (8, 64, 7)         # cfg 2 / cfg 4
    assert geom(20000, 1000, 0) == geom(20000, 1000, 1) == (79, 256, 16)                # >= 8 chunks already: unchanged
    assert geom(72, 72, 1) == (8, 16, 2) and geom(17, 8, 1) == (8, 16, 1)               # the narrowest UMMA
    for N in (1, 15, 16, 17, 100, 255, 256, 257, 511, 513, 1000, 2048, 3125, 8582):
        for split in (0, 1):
            nch, nc, kb = geom(N, 77, split)
            assert nc % 16 == 0 and 16 <= nc <= 256 and nch * nc >= N and kb == 2
            assert nch >= (8 if split else 1)
            if not split:
                assert (nch - 1) * nc < N       # (the 16-column granularity can leave trailing all-padding chunks in the split geometry)
        assert geom(N, 77, 1)[0] >= geom(N, 77, 0)[0]
    a = C.c_int()
    assert lib.sdrm_layer_geometry(0, 4, 0, C.byref(a), C.byref(a), C.byref(a)) != 0
