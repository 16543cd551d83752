"""TEST INFRASTRUCTURE — import the UNMODIFIED reference from /root/reference (CPU only).

/root/reference exists only in the build container; it does NOT exist on the GPU box.
Nothing that runs there (gpu tests, smoke(), bench.py) may call this module.  It is
used by ``oracle/make_golden.py`` and by the ``not gpu`` tests that pin
``oracle/tcdiff_oracle.py`` against the real thing (they skip when the tree is absent).

The reference imports render-only / absent packages at module import time
(model/diffusion.py:12-18, vis.py:5-17).  We register inert in-memory stand-ins for
them and a pure-torch restatement (``oracle/p3d.py``) for ``pytorch3d.transforms``.
No reference file is modified or copied.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TCDIFF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "diffusion.py"))


def _inert(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda key: (_ for _ in ()).throw(AttributeError(key)) if key.startswith("__") else _Dummy()
    sys.modules[name] = m
    return m


class _Dummy:
    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, k):
        return _Dummy()


_loaded = {}


def load():
    """Returns a namespace with the reference's DanceDecoder, GaussianDiffusion, SMPLSkeleton,
    RotaryEmbedding, ax_from_6v, and the imported modules themselves."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import p3d

    if "pytorch3d" not in sys.modules:
        pk = types.ModuleType("pytorch3d")
        tr = types.ModuleType("pytorch3d.transforms")
        for k in dir(p3d):
            if not k.startswith("_") and callable(getattr(p3d, k)):
                setattr(tr, k, getattr(p3d, k))
        pk.transforms = tr
        sys.modules["pytorch3d"] = pk
        sys.modules["pytorch3d.transforms"] = tr
    for name in ["p_tqdm", "librosa", "soundfile", "matplotlib", "matplotlib.animation",
                 "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _inert(name)
    if "p_tqdm" in sys.modules and not hasattr(sys.modules["p_tqdm"], "p_map"):
        sys.modules["p_tqdm"].p_map = lambda f, *its, **k: [f(*a) for a in zip(*its)]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's top-level package names are generic ("model", "dataset", "vis")
    for clash in ("model", "dataset", "vis"):
        mod = sys.modules.get(clash)
        if mod is not None and REFERENCE_ROOT not in (getattr(mod, "__file__", "") or ""):
            del sys.modules[clash]
    model_model = importlib.import_module("model.model")
    model_diff = importlib.import_module("model.diffusion")
    model_rot = importlib.import_module("model.rotary_embedding_torch")
    model_utils = importlib.import_module("model.utils")
    vis = importlib.import_module("vis")
    quat = importlib.import_module("dataset.quaternion")
    # silence tqdm bars in the reference loops
    model_diff.tqdm = lambda it, **k: it
    ns = types.SimpleNamespace(
        DanceDecoder=model_model.DanceDecoder,
        GaussianDiffusion=model_diff.GaussianDiffusion,
        SMPLSkeleton=vis.SMPLSkeleton,
        RotaryEmbedding=model_rot.RotaryEmbedding,
        ax_from_6v=quat.ax_from_6v,
        model_model=model_model, model_diffusion=model_diff, model_utils=model_utils,
        model_rotary=model_rot, vis=vis, quaternion=quat,
    )
    _loaded["ns"] = ns
    return ns


class NoiseBank:
    """Context manager: make the reference consume pre-drawn noise.

    The reference draws noise internally (model/diffusion.py:393,421,246,269,644) and the CFG
    keep mask in model/model.py:567.  Inside this context ``torch.randn`` / ``torch.randn_like``
    pop tensors from ``bank`` (in call order) and ``prob_mask_like`` returns ``keep_mask`` when
    0 < prob < 1.
    """

    def __init__(self, bank, keep_mask=None):
        self.bank = list(bank)
        self.keep_mask = keep_mask
        self.i = 0

    def _pop(self, shape):
        import torch
        t = self.bank[self.i]
        self.i += 1
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.clone()

    def __enter__(self):
        import torch
        ns = load()
        self._randn, self._randn_like = torch.randn, torch.randn_like
        self._pml = ns.model_model.prob_mask_like
        torch.randn = lambda *shape, **k: self._pop(shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape)
        torch.randn_like = lambda x, **k: self._pop(x.shape)
        km = self.keep_mask

        def pml(shape, prob, device):
            if km is not None and 0 < prob < 1:
                return km.to(device)
            return self._pml(shape, prob, device)
        ns.model_model.prob_mask_like = pml
        return self

    def __exit__(self, *exc):
        import torch
        torch.randn, torch.randn_like = self._randn, self._randn_like
        load().model_model.prob_mask_like = self._pml
        return False
