"""Test helpers: the scene -> object graph -> flat C-ABI arrays construction lives in the package (bundle_adjustment_b200.workloads)."""
from bundle_adjustment_b200.workloads import FIXED, build_adjustment, flat_problem  # noqa: F401
