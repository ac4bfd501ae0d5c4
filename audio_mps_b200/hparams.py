"""Minimal stand-in for ``tf.contrib.training.HParams`` as the reference uses it
(/root/reference/train.py:41-44, tests/test_model.py:13-14): keyword construction, attribute
access and ``parse("name=value,...")`` overrides typed after the existing value."""
from __future__ import annotations


class HParams:
    def __init__(self, **kwargs):
        self._names = []
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._names:
            raise ValueError(f"Hyperparameter name is reserved or already defined: {name}")
        self._names.append(name)
        setattr(self, name, value)

    def set_hparam(self, name, value):
        if name not in self._names:
            raise ValueError(f"Unknown hyperparameter: {name}")
        setattr(self, name, value)

    def parse(self, values: str):
        """Override from a comma separated ``name=value`` list (train.py:44)."""
        if not values:
            return self
        for item in values.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" not in item:
                raise ValueError(f"Could not parse hparam assignment: {item!r}")
            name, raw = item.split("=", 1)
            name = name.strip()
            if name not in self._names:
                raise ValueError(f"Unknown hyperparameter: {name}")
            cur = getattr(self, name)
            raw = raw.strip()
            if isinstance(cur, bool):
                val = raw.lower() in ("1", "true", "t", "yes")
            elif isinstance(cur, int):
                val = int(raw)
            elif isinstance(cur, float):
                val = float(raw)
            elif cur is None:
                try:
                    val = int(raw)
                except ValueError:
                    try:
                        val = float(raw)
                    except ValueError:
                        val = None if raw.lower() == "none" else raw
            else:
                val = type(cur)(raw)
            setattr(self, name, val)
        return self

    def values(self):
        return {k: getattr(self, k) for k in self._names}

    def __contains__(self, name):
        return name in self._names

    def __repr__(self):
        return "HParams(" + ", ".join(f"{k}={getattr(self, k)!r}" for k in self._names) + ")"


def default_hparams(sample_rate: int = 16000) -> HParams:
    """The hparams ``train.py`` builds (train.py:41-43)."""
    import math
    return HParams(minibatch_size=8, bond_dim=8, delta_t=1 / sample_rate, sigma=0.0001,
                   h_reg=200 / (math.pi * sample_rate) ** 2, r_reg=0.1,
                   initial_rank=None, A=100., learning_rate=0.001)
