"""Reference-side bindings of the C ABI (what a sitator maintainer would add); quoted by INTEGRATION.md and executed
by tests/test_reference_binding_gpu.py against the compiled reference."""
