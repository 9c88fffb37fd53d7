"""TEST INFRASTRUCTURE ONLY -- stand-in for TensorFlow: the reference's network module evaluates
tf.keras.optimizers.Adam() as a default argument at import time (network.py:78); nothing on the decoding path uses it."""
import sys
import types


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self


sys.modules[__name__] = _Anything(__name__)
