"""Harness side of the bench contract: `baseline/_ref/` (git-ignored) holds what was taken
UNMODIFIED from the reference at build time (oracle/build_ref.py); the tracked files here only
load it.  Not product code: ganet_b200/ never imports this package."""
