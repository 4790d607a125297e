"""Import-only stand-in: models/__init__.py imports the classical Wavelet baseline, which imports pywt.  Nothing on the
LGTEUN path calls it."""


def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(f"pywt.{name}: PyWavelets is not installed (shims/pywt.py is an import stub)")
    return _missing
