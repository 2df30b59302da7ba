"""Import alias: the package directory is `whisper-vits-svc_b200/` (not a valid Python
identifier), so `import whisper_vits_svc_b200` resolves to it through this shim."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "whisper-vits-svc_b200")]
with open(_os.path.join(__path__[0], "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
