"""Import the UNMODIFIED reference (read-only, /root/reference) inside the build container.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call `load_reference()`; `available()` is the guard.

The only obstacle to importing GPzoo here is a top-level `import matplotlib.pyplot`
(gpzoo/utilities.py:11) — matplotlib is not installed — so an empty stub module is injected
into sys.modules first.  No reference file is copied or modified.
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("GPZOO_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "gpzoo"))


def load_reference():
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import gpzoo.gp as gp            # noqa: E402
    import gpzoo.kernels as kernels  # noqa: E402
    import gpzoo.likelihoods as lik  # noqa: E402
    import gpzoo.utilities as util   # noqa: E402
    return types.SimpleNamespace(gp=gp, kernels=kernels, likelihoods=lik, utilities=util)


@contextlib.contextmanager
def fixed_eps(*eps_list):
    """Make Normal.rsample (torch normal.py) consume the given standard-normal draws, in order,
    instead of the global RNG, so the reference and the CUDA path see the same `eps`."""
    import torch.distributions.normal as tn
    queue = list(eps_list)
    orig = tn._standard_normal

    def fake(shape, dtype, device):
        e = queue.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype=dtype, device=device)

    tn._standard_normal = fake
    try:
        yield
    finally:
        tn._standard_normal = orig


@contextlib.contextmanager
def quiet():
    """VNNGP.forward prints unconditionally (gp.py:32,65,79-81,84)."""
    with open(os.devnull, "w") as f, contextlib.redirect_stdout(f):
        yield


@contextlib.contextmanager
def exact_cdist(flag=True):
    """Force torch.cdist onto its direct-difference path inside the reference process."""
    orig = torch.cdist
    if flag:
        def patched(x1, x2, p=2.0, compute_mode="use_mm_for_euclid_dist_if_necessary"):
            return orig(x1, x2, p=p, compute_mode="donot_use_mm_for_euclid_dist")
        torch.cdist = patched
    try:
        yield
    finally:
        torch.cdist = orig
