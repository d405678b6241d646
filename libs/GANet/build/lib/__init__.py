"""`from libs.GANet.build.lib import GANet` (reference functions/GANet.py:3) resolves to
the legacy-named native entry points implemented over the C ABI."""
from ganet_b200 import legacy_native as GANet  # noqa: F401
