"""`hp` loading without OmegaConf.

The reference reads configs/base.yaml with OmegaConf and uses attribute access
(`hp.vits.* / hp.gen.* / hp.data.*`, svc_inference.py:162-166).  OmegaConf is an
optional dependency here: a plain yaml.safe_load into an attribute dict exposes the
same surface; an OmegaConf DictConfig passed by a caller works unchanged.
"""
from __future__ import annotations

import copy

import yaml


class HParams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return HParams({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_hparams(d):
    if isinstance(d, dict):
        return HParams({k: to_hparams(v) for k, v in d.items()})
    return d


def load_hparams(path: str) -> HParams:
    with open(path, "r", encoding="utf-8") as f:
        return to_hparams(yaml.safe_load(f))


def override(hp: HParams, **dotted) -> HParams:
    """Return a deep copy with `section__key=value` overrides, e.g.
    override(hp, gen__upsample_input=80, data__sampling_rate=24000) — BASELINE config #2."""
    out = copy.deepcopy(hp)
    for k, v in dotted.items():
        sec, key = k.split("__", 1)
        out[sec][key] = v
    return out
