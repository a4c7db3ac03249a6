"""Compiled pybind11 torch-extension modules (built in-tree by pienerf_b200/build_ext.py); see pienerf_b200.dropin."""
