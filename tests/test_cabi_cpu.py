"""CPU: the C-ABI library builds, loads and exports every symbol include/gpzoo_b200.h declares; the product
path refuses CPU tensors loudly (there is no fallback)."""
import os

import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    from gpzoo_b200 import _cabi, build
    if not os.path.exists(_cabi.LIB_PATH):
        build.build()
    return _cabi.lib()


def test_exports_match_header(lib):
    from gpzoo_b200 import _cabi
    syms = _cabi.exported_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.gpz_abi_version() >= 1
    assert b"success" in lib.gpz_error_string(0)


def test_no_cpu_fallback(lib):
    import gpzoo_b200 as gz
    k = gz.kernels.NSF_RBF(L=2)
    with pytest.raises(gz._cabi.GpzError):
        k(torch.randn(5, 2), torch.randn(3, 2))


def test_initialisation_has_no_host_path():
    """The initialisation pipeline refuses to run without a CUDA device; its pure formulas need none."""
    import numpy as np
    from gpzoo_b200 import initialisation as I, utilities as U
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            I.nmf(np.ones((6, 4)), 2)
        with pytest.raises(RuntimeError):
            U.regularized_nmf(np.ones((6, 4)), 2)
    mu, sd = U.lnormal_approx_dirichlet(4)
    assert abs(mu - (-np.log(4) - (np.log(8) - np.log(5)) / 2)) < 1e-15 and abs(sd - np.sqrt(np.log(8) - np.log(5))) < 1e-15
    assert set(("regularized_nmf", "shrink_factors", "shrink_loadings", "init_softplus", "rescale_spatial_coords",
                "build_group_distances", "lnormal_approx_dirichlet")) <= set(dir(U))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "gpzoo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_fused_adam_and_tiled_predict_refuse_cpu(lib):
    """The pieces next to the path have no CPU route either: the fused optimiser and the tiled predictor raise on CPU tensors."""
    import gpzoo_b200 as gz
    p = torch.nn.Parameter(torch.randn(7))
    p.grad = torch.randn(7)
    with pytest.raises(gz._cabi.GpzError):
        gz.optim.Adam([p], lr=1e-2).step()
    gp = gz.gp.SVGP(gz.kernels.NSF_RBF(L=2), dim=2, M=8, jitter=1e-2)
    gp.mu = torch.nn.Parameter(torch.zeros(2, 8))
    gp.Lu = torch.nn.Parameter(torch.eye(8).expand(2, -1, -1).contiguous())
    with pytest.raises(gz._cabi.GpzError):
        gp.predict_moments(torch.randn(20, 2), tile=8)


def test_abi_version_and_new_entry_points(lib):
    """ABI 2 added `kind` to the kernel-build calls; the split-FP16 / large-M / training-step entry points are exported."""
    assert lib.gpz_abi_version() >= 2
    for s in ("gpz_kernel_build_fwd_h_f32", "gpz_svgp_predict_fwd_h_f32", "gpz_svgp_predict_bwd_h_f32", "gpz_split16_f32",
              "gpz_umma_gemm16_f32", "gpz_chol_inv_tc_f32", "gpz_adam_step_f32", "gpz_adam_step_f64", "gpz_svgp_predict_h_stat_row"):
        assert hasattr(lib, s), s
    assert lib.gpz_svgp_predict_h_supported(1024, 32768) == 1 and lib.gpz_svgp_predict_h_supported(1020, 32768) == 0
    rows = [lib.gpz_svgp_predict_h_stat_row(i) for i in range(5)]
    assert len(set(rows)) == 5 and min(rows) >= 0                       # five distinct rows of the stats block
    assert lib.gpz_svgp_predict_h_stat_row(99) == -1


def test_state_dict_keys_and_shapes_match_reference():
    """Every model family has the reference's state_dict keys and shapes (golden written from the unmodified reference by
    oracle/gen_golden.py; compared live as well when /root/reference is present).  Module construction needs no GPU."""
    import json
    import os
    import types
    import gpzoo_b200 as gz
    from oracle import ref_loader, ref_runner
    from tests.helpers import GOLDEN
    golden = json.load(open(os.path.join(GOLDEN, "state_dict_shapes.json")))
    ours = ref_runner.state_dict_shapes(types.SimpleNamespace(kernels=gz.kernels, gp=gz.gp, likelihoods=gz.likelihoods))
    assert set(ours) == set(golden)
    for name in golden:
        assert ours[name] == golden[name], (name, ours[name], golden[name])
    if ref_loader.available():
        assert ref_runner.state_dict_shapes() == golden


def test_reference_utility_names_importable():
    """`from gpzoo.utilities import ...` of every hot-path name keeps working after install_as_gpzoo()."""
    import sys
    import gpzoo_b200 as gz
    saved = {k: v for k, v in sys.modules.items() if k == "gpzoo" or k.startswith("gpzoo.")}
    gz.install_as_gpzoo()
    try:
        _check_reference_names()
    finally:                                  # other tests import the real reference under the same name
        for k in [k for k in sys.modules if k == "gpzoo" or k.startswith("gpzoo.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _check_reference_names():
    from gpzoo.utilities import (_embed_distance_matrix, _squared_dist, _torch_sqrt, add_jitter, reshape_param, svgp_forward,  # noqa: F401
                                 train, train_batched, train_closure_batched, train_hybrid, train_hybrid_batched, whitened_KL)
    from gpzoo.gp import MGGP_WSVGP, WSVGP
    from gpzoo.likelihoods import Hybrid_NSF2
    assert hasattr(WSVGP, "forward_precomputed") and hasattr(MGGP_WSVGP, "forward_precomputed")
    assert hasattr(Hybrid_NSF2, "forward_precomputed")
    import torch
    assert reshape_param(torch.zeros(2, 3, 4, 5)).shape == (6, 4, 5)
    assert abs(float(_torch_sqrt(torch.zeros((), dtype=torch.float64), 1e-6)) - 1e-3) < 1e-15
