"""`import tensor_resize` shim: put pats_b200/shims on sys.path (or call pats_b200.install.install())
and utils/utils.py:17 of the reference picks up the CUDA implementation instead of the compiled
setup/library.cpp."""
from pats_b200.tensor_resize import tensor_resize  # noqa: F401
