"""Import-path shim: ``from digicamtoy.tracegenerator import NTraceGenerator`` (the reference's
import, e.g. produce_data.py:5, example.py:1) resolves to the B200 implementation."""
