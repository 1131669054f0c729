class _Missing:
    def __getattr__(self, name):
        raise ImportError("astropy is not installed; this is an import shim for the oracle")


fits = _Missing()
ascii = _Missing()
