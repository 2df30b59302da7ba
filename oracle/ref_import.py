"""Import shim for the *unmodified* reference tree (TEST INFRASTRUCTURE).

Only usable where the reference checkout exists (the build container:
/root/reference, or $SVCB_REFERENCE).  Nothing that runs on the GPU box may
depend on it; tests that need it skip when it is absent.

* `hp` stand-in: the reference reads its YAML with OmegaConf (svc_inference.py:162)
  which is not installed here -> yaml.safe_load into an attribute dict.
* `whisper.model` imports `.decoding` -> `.audio` -> `import librosa`
  (whisper/model.py:11, whisper/audio.py:5): a stub module with a valid
  `__spec__` is injected first.
* RNG capture: `record_rng()` wraps torch.randn_like / torch.rand so the draws the
  reference makes (vits/models.py:51, vits_decoder/nsf.py:232-236,311) can be fed to
  the oracle and the CUDA path.
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("SVCB_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "vits", "models.py"))


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


def load_hp(path: str | None = None):
    path = path or os.path.join(REF_ROOT, "configs", "base.yaml")
    with open(path) as f:
        return to_attr(yaml.safe_load(f))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def _stub(name):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    sys.modules[name] = m


def import_synthesizer():
    _ensure_path()
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)
    from vits.models import SynthesizerInfer  # noqa
    return SynthesizerInfer


def import_whisper_model():
    _ensure_path()
    _stub("librosa")
    _stub("librosa.filters")
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    if not hasattr(sys.modules["librosa.filters"], "mel"):
        def _no_mel(*a, **k):
            raise RuntimeError("librosa is not installed; the log-mel front-end is outside the oracle's scope")
        sys.modules["librosa.filters"].mel = _no_mel
    import whisper.model as wm  # noqa
    return wm


def import_hubert_model():
    """hubert/hubert_model.py (needs only torch)."""
    _ensure_path()
    import hubert.hubert_model as hm  # noqa
    return hm


@contextlib.contextmanager
def record_rng(log: list):
    """Record every tensor returned by torch.randn_like / torch.rand inside the block."""
    o_randn_like, o_rand = torch.randn_like, torch.rand

    def randn_like(*a, **k):
        t = o_randn_like(*a, **k)
        log.append(("randn_like", t.clone()))
        return t

    def rand(*a, **k):
        t = o_rand(*a, **k)
        log.append(("rand", t.clone()))
        return t

    torch.randn_like, torch.rand = randn_like, rand
    try:
        yield log
    finally:
        torch.randn_like, torch.rand = o_randn_like, o_rand
