"""Physical constants shared with the reference (values must match bit-for-bit)."""

#: speed of light in km/s — same literal as the reference (Starfish/constants.py:7)
c_kms = 2.99792458e5

#: diagonal jitter the likelihood adds before factorising (Starfish/models/spectrum_model.py:399)
JITTER = 1e-10
