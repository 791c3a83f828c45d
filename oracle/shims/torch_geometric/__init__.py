"""Minimal stand-in for torch_geometric (absent from this image) so that the UNMODIFIED
reference modules under /root/reference import and run on CPU.  TEST INFRASTRUCTURE ONLY:
nothing in the product package imports this."""
