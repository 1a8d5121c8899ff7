"""The example configurations live in the package (bulletproofspp_b200/workloads.py) so that bench.py and
tools/ do not import from tests/; re-exported here for the tests."""
from bulletproofspp_b200.workloads import EXAMPLES, batched  # noqa: F401
