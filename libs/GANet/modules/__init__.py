from .GANet import *  # noqa: F401,F403  (same as the reference's modules/__init__.py)
