"""pienerf_b200 — B200-native (sm_100a) sim+render hot path of PIE-NeRF behind the reference's operator API.

Importing the package loads pienerf_b200/lib/libpienerf_b200.so and fails loudly if it is missing; there is
no CPU fallback.  `synthetic` is importable without the library (fixtures only)."""
__version__ = "0.1.0"
