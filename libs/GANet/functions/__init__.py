from .GANet import *  # noqa: F401,F403
