"""The C-ABI library loads and exports every symbol include/svcb.h declares (no compute)."""
import os
import re

import pytest

from tests.util import ROOT
from whisper_vits_svc_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "svcb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svcb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert "svcb_infer" in names and "svcb_model_create" in names
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} not exported by libsvc_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.svcb_version() >= 100
    assert isinstance(lib.svcb_last_error(), (bytes, type(None)))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SvcbError):
        _lib.load()


def test_no_cpu_path():
    import torch
    from whisper_vits_svc_b200 import hparams, models
    hp = hparams.load_hparams(os.path.join(ROOT, "configs", "base.yaml"))
    m = models.SynthesizerInfer(513, 25, hp)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SvcbError):
            m.pitch2source(torch.zeros(1, 4))


def test_config_struct_layout_matches_header():
    import ctypes
    lib = _lib.load()
    assert ctypes.sizeof(_lib.Config) == lib.svcb_sizeof(0) == 4 * 52
    assert ctypes.sizeof(_lib.TensorEntry) == lib.svcb_sizeof(1) == 96 + 16
    assert ctypes.sizeof(_lib.Taps) == lib.svcb_sizeof(2) == 8 * _lib.SVCB_TAP_COUNT
