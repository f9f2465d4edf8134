from digicamtoy_b200.tracegenerator import NTraceGenerator  # noqa: F401
