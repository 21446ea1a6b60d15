"""Pure-torch stand-in for the parts of `torch_geometric` 2.0 the reference layer imports
(TEST INFRASTRUCTURE ONLY; see oracle/reference_loader.py)."""
__version__ = "2.0.0-oracle-shim"
