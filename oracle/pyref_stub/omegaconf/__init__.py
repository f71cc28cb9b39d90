"""Minimal stand-in for `omegaconf` so the UNMODIFIED reference package under
/root/reference can be imported in this container (hydra/omegaconf are absent).

TEST INFRASTRUCTURE ONLY (oracle/): used by oracle/ref_harness.py to generate
golden vectors.  The reference only uses `DictConfig` as a type hint and as an
attribute-access mapping (e.g. dpLGAR/models/dpLGAR.py:41-72).
"""


class DictConfig(dict):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, DictConfig):
                self[k] = DictConfig(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, DictConfig):
            v = DictConfig(v)
        self[k] = v
